#!/bin/bash
# artefacts for profiles/: launch list of the bench command + one --set full capture of each hot kernel
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_raw.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 0.05 --e2e-steps 1 > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_raw.csv > gpurun_out/${TAG}_launches.txt; cat gpurun_out/${TAG}_launches.txt
for k in k_tile_pass k_wide_scatter k_ffat_update_buckets; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 8 -c 1 -f -o gpurun_out/${TAG}_$k python bench.py --steps 2 --warmup 3 --cpu-seconds 0.05 --e2e-steps 1 > gpurun_out/${TAG}_$k.log 2>&1
  echo "$k rc=$?"
done
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json
