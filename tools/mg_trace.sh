#!/bin/bash
# N GPUs: the checked bench run with the device timeline of the step (WFB_MG_TRACE=1: the library prints it every 64 steps); `nccl`: also the NCCL exchange (WFB_MG_CE=0)
N=${1:-2}
mkdir -p gpurun_out
run() { tag=$1; shift
env "$@" WFB_MG_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps ${STEPS:-130} --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 --no-extras > gpurun_out/mg_${N}_$tag.json 2> gpurun_out/mg_${N}_$tag.err
grep "wfb_mg rank" gpurun_out/mg_${N}_$tag.err | grep -v "us/step" | head -1 | cut -c1-150
grep "wfb_mg rank 0" gpurun_out/mg_${N}_$tag.err | tail -1 | cut -c1-220
python -c "
import json
for l in open('gpurun_out/mg_${N}_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N $tag', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e9,3), 'win', d['config']['windows_in_timed_region'], d['config']['windows_expected_steady_state'], 'check', d['check'] and (d['check']['passed'], d['check']['windows_compared']))
" || tail -5 gpurun_out/mg_${N}_$tag.err | cut -c1-300
}
run ${TAG:-ce} ${EXTRA_ENV}
if [ "$2" == "nccl" ]; then run nccl WFB_MG_CE=0; fi
