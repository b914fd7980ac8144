#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
for items in 4 8 16; do for bps in 64 256; do
  WFB_OS_ITEMS=$items timeout 200 python bench.py --steps 20 --warmup 3 --batches-per-step $bps --cpu-seconds 0.2 --e2e-steps 2 > gpurun_out/exp_${items}_${bps}.json 2>gpurun_out/exp.err
  python -c "
import json; d=json.load(open('gpurun_out/exp_${items}_${bps}.json')); p=d['roofline']['phase_ms_per_step']; print('items=$items bps=$bps', round(d['value']/1e9,2),'GT/s', {k:round(v,3) for k,v in p.items()}, 'launches', d['gpu_launches'])"
done; done
