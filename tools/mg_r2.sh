#!/bin/bash
# multi-GPU check of round 2: GPU suite subset on rank 0, then the bench at N GPUs (with the full-size check)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_keyed.py tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/mg_pytest.log 2>&1; echo "pytest rc=$? $(grep -E 'passed|failed' gpurun_out/mg_pytest.log | tail -1)"
for mode in "" "--sync-exchange"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-30} --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 --no-extras $mode > gpurun_out/mg_$N$mode.json 2> gpurun_out/mg_$N$mode.err
python -c "
import json,sys
for l in open('gpurun_out/mg_$N$mode.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N $mode', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']/1e9,3), 'win', d['config']['windows_in_timed_region'], d['config']['windows_expected_steady_state'], 'check', d['check'])
" || tail -5 gpurun_out/mg_$N$mode.err
done
