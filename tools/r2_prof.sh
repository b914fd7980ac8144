#!/bin/bash
mkdir -p gpurun_out
python tools/torch_overhead.py 2>&1 | tail -9
B() { tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 40 --warmup 5 --cpu-seconds 0.2 --e2e-steps 2 --no-extras --no-check > gpurun_out/exp_$tag.json 2>gpurun_out/exp_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/exp_$tag.json')); k=d['roofline']['kernels']; print('$tag', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), [round(x['avg_us'],1) for x in k], d['gpu_launches'])" || tail -5 gpurun_out/exp_$tag.err
}
B default
B bk4 WFB_LAZY_TREE=0
bash tools/prof_kernel.sh k_wide_scatter_ranked r2f_scatter_ranked
bash tools/prof_kernel.sh k_ffat_update_buckets r2f_update_buckets
for K in 128 64 16 4 1; do timeout 300 windflow_b200/apps/pipeline_bench.bin $K $((K>8?20480:2048)) 2>&1 | tail -1; done
