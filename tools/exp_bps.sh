#!/bin/bash
mkdir -p gpurun_out
for bps in 16 32 64 128 256; do
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.2 --e2e-steps 2 --batches-per-step $bps > gpurun_out/exp_bps$bps.json 2>gpurun_out/exp_bps$bps.err
  python -c "
import json; d=json.load(open('gpurun_out/exp_bps$bps.json')); p=d['roofline']['phase_ms_per_step']; print('bps=$bps', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],3), {k:round(v,3) for k,v in p.items()})" || tail -3 gpurun_out/exp_bps$bps.err
done
