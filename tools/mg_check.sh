#!/bin/bash
# multi-GPU path: tests on one GPU (world 1), then the bench at the box's GPU count
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_keyed.py tests/test_gpu_ffat.py -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
N=${1:-2}
for mode in "" "--sync-exchange"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 $mode > gpurun_out/mg_$N$mode.json 2> gpurun_out/mg_$N$mode.err
python -c "
import json,sys
for l in open('gpurun_out/mg_$N$mode.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N $mode', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],3), d['roofline'].get('phase_ms_per_step'), 'e2e', round(d['e2e']['value']/1e9,3))
" || tail -5 gpurun_out/mg_$N$mode.err
done
