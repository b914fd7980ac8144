#!/bin/bash
# round-2 experiment driver: GPU suite on the default / forced-stream / old bucket paths, then bench variants (per-phase times)
mkdir -p gpurun_out
T() { tag=$1; shift; env "$@" timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest[$tag] rc=$?  $(grep -E 'passed|failed|error' gpurun_out/pytest_$tag.log | tail -1)"; }
B() { tag=$1; shift
  env "$@" timeout 600 python bench.py --steps ${STEPS:-40} --warmup 5 --cpu-seconds 0.2 --e2e-steps 2 --no-extras $BARGS > gpurun_out/exp_$tag.json 2>gpurun_out/exp_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/exp_$tag.json')); k=d['roofline']['kernels']; print('$tag', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), [round(x['avg_us'],1) for x in k], d['gpu_launches'], 'host', round(d['host_issue_ms_per_step'],3), 'win', d['config']['windows_in_timed_region'], d['config']['windows_expected_steady_state'], 'check', d.get('check') and d['check']['windows_compared'])" || tail -5 gpurun_out/exp_$tag.err
}
if [ "$1" != "nobench" ] && [ "$1" != "notest" ]; then
T default
T stream WFB_UPDATE=stream
T old WFB_UPDATE=buckets WFB_TILE_H16=0
fi
B default
BARGS="--no-check" B stream WFB_UPDATE=stream
BARGS="--no-check --pipeline" B pipelined
BARGS="--no-check --pipeline" B pipelined_ctas1 WFB_INGEST_CTAS_PER_SM=1
