#!/bin/bash
# round-2 experiment driver: GPU suite on the default / forced-stream / old bucket paths, then bench variants (per-phase times)
mkdir -p gpurun_out
T() { tag=$1; shift; env "$@" timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest[$tag] rc=$?  $(tail -1 gpurun_out/pytest_$tag.log)"; }
B() { tag=$1; shift
  env "$@" timeout 300 python bench.py --batches-per-step ${BPS:-128} --steps ${STEPS:-40} --warmup 5 --cpu-seconds 0.2 --e2e-steps 2 > gpurun_out/exp_$tag.json 2>gpurun_out/exp_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/exp_$tag.json')); p=d['roofline']['phase_ms_per_step']; print('$tag', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],4), {k:round(v,4) for k,v in p.items()}, d['gpu_launches'])" || tail -5 gpurun_out/exp_$tag.err
}
if [ "$1" != "nobench" ]; then
T default
T stream WFB_UPDATE=stream
T old WFB_UPDATE=buckets WFB_TILE_H16=0
fi
B default
B old WFB_UPDATE=buckets WFB_TILE_H16=0
B h16_buckets WFB_UPDATE=buckets
B bps64 BPS=64
if [ -x oracle/_ref/t_fat_gpu_tb ]; then timeout 300 oracle/_ref/t_fat_gpu_tb -r 3 -l 200000 -k 13 -w 5000 -s 1000 2>&1 | grep -v "^|\|^+" | grep -E "Result|threads|Error|error" | head; echo "ref gpu test rc=$?"; fi
if [ -x oracle/_ref/ref_pipeline_gpu ]; then
  timeout 300 oracle/_ref/ref_pipeline_gpu gpu_cb gen=1048576 keys=64 batch=65536 win=4096 slide=64 nb=65 out=gpurun_out/ref_gpu_cb.out 2>&1 | tail -1
  timeout 300 oracle/_ref/ref_pipeline_gpu gpu_cb gen=2097152 keys=65536 batch=65536 win=4096 slide=64 nb=65 2>&1 | tail -1
  timeout 300 oracle/_ref/ref_pipeline_gpu gpu_mf gen=4194304 batch=65536 2>&1 | tail -1
  timeout 300 oracle/_ref/ref_pipeline_cpu cpu_cb gen=4194304 keys=65536 win=4096 slide=64 par=16 2>&1 | tail -1
fi
