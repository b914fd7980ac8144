#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ffat.py tests/test_gpu_ffat_tb.py tests/test_gpu_keyed.py -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.2 --e2e-steps 2 > gpurun_out/exp_one.json 2>gpurun_out/exp_one.err
python -c "
import json; d=json.load(open('gpurun_out/exp_one.json')); p=d['roofline']['phase_ms_per_step']; print(round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],3), {k:round(v,3) for k,v in p.items()}, d['gpu_launches'], 'frac', round(d['roofline']['frac'],3))" || tail -5 gpurun_out/exp_one.err
