#!/bin/bash
# multi-GPU check of round 2: the bench at N GPUs (with the full-size check): bucketed exchange (default), two-partition exchange, Python-driven exchange
N=${1:-2}
mkdir -p gpurun_out
run() { tag=$1; shift
env $ENVV timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-65} --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 --no-extras "$@" > gpurun_out/mg_${N}_$tag.json 2> gpurun_out/mg_${N}_$tag.err
python -c "
import json,sys
for l in open('gpurun_out/mg_${N}_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N $tag', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],3), 'host', round(d['host_issue_ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e9,3), 'win', d['config']['windows_in_timed_region'], d['config']['windows_expected_steady_state'], 'check', d['check'])
" || tail -5 gpurun_out/mg_${N}_$tag.err
}
run bucketed
if [ "$2" != "only" ]; then
ENVV="WFB_MG_BUCKETED=0" run twopart
run pyexchange --py-exchange
fi
