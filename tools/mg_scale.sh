#!/bin/bash
# scaling run on one box: bench at the given GPU counts (default 2 4 8)
mkdir -p gpurun_out
for N in ${@:-2 4 8}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 30 --warmup 5 --cpu-seconds 0.2 --e2e-steps 2 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
python -c "
import json
for l in open('gpurun_out/scale_$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline'].get('phase_ms_per_step',{}).items()}, 'e2e', round(d['e2e']['value']/1e9,3))
" || tail -5 gpurun_out/scale_$N.err
done
