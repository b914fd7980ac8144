#!/bin/bash
# The reference's own GPU kernels + Thrust sequences on this box: ncu launch lists of oracle/_ref/ref_pipeline_gpu (the unmodified
# wf/windflow_gpu.hpp operators) for BASELINE cfg 2 (Map_GPU -> Filter_GPU), cfg 3 (Reduce_GPU keyed) and cfg 4 (Ffat_Windows_GPU CB, 64 and 4096 keys).
# Per-launch times are cold-cache and serialised; the sum per batch is the reference's device time per 65536-tuple batch.
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/ref_${tag}_raw.csv oracle/_ref/ref_pipeline_gpu "$@" > gpurun_out/ref_${tag}.log 2>&1
  echo "== $tag: $(tail -1 gpurun_out/ref_${tag}.log)"; python tools/launch_summary.py gpurun_out/ref_${tag}_raw.csv | tee gpurun_out/ref_${tag}_launches.txt
}
run mf gpu_mf gen=524288 batch=65536
run red gpu_red gen=524288 keys=1000000 batch=65536
run cb64 gpu_cb gen=1048576 keys=64 batch=65536 win=4096 slide=64 nb=65
run cb4096 gpu_cb gen=262144 keys=4096 batch=65536 win=4096 slide=64 nb=65
