#!/bin/bash
# N GPUs: the bench with the device timeline of the step (WFB_MG_TRACE), then the checked run; optionally the NCCL exchange for comparison
N=${1:-2}
env WFB_MG_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 130 --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 --no-extras --no-check > gpurun_out/mg_${N}_trace.json 2> gpurun_out/mg_${N}_trace.err
grep "wfb_mg rank" gpurun_out/mg_${N}_trace.err | head -2 | cut -c1-150
grep "wfb_mg rank 0" gpurun_out/mg_${N}_trace.err | tail -1 | cut -c1-220
tail -2 gpurun_out/mg_${N}_trace.err | cut -c1-300
bash tools/mg_r2.sh $N only 2>&1 | cut -c1-200
if [ "$2" == "nccl" ]; then
env WFB_MG_CE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 65 --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 --no-extras --no-check > gpurun_out/mg_${N}_nccl.json 2> gpurun_out/mg_${N}_nccl.err
python -c "
import json
for l in open('gpurun_out/mg_${N}_nccl.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N nccl exchange', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],3))
"
fi
