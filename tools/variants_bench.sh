#!/bin/bash
# bench.py (N=1, no check) with the default library and with every build under windflow_b200/variants/ (compile-time knobs, loaded through WFB_LIB)
mkdir -p gpurun_out
B() { tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 65 --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 --no-extras --no-check > gpurun_out/exp_$tag.json 2>gpurun_out/exp_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/exp_$tag.json')); k=d['roofline']['kernels']; print('$tag', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), [round(x['avg_us'],1) for x in k])" || tail -5 gpurun_out/exp_$tag.err
}
B default
for v in windflow_b200/variants/*.so; do B $(basename $v .so) WFB_LIB=$PWD/$v; done
