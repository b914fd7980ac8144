#!/bin/bash
# bench every build under windflow_b200/variants/ against the default library
mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 $EXTRA > gpurun_out/exp_$tag.json 2>gpurun_out/exp_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/exp_$tag.json')); p=d['roofline'].get('phase_ms_per_step',{}); print('$tag', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],3), {k:round(v,3) for k,v in p.items()}, d['gpu_launches'])" || tail -5 gpurun_out/exp_$tag.err
}
run base
for v in windflow_b200/variants/*.so; do run $(basename $v .so) WFB_LIB=$PWD/$v; done
