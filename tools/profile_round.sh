#!/bin/bash
# artefacts for profiles/: launch list of the bench command (N=1 path, and the N>1 step with one rank) + one --set full capture of each hot
# kernel + the facade's launch list; the .ncu-rep files are summarised by tools/profile_summaries.sh afterwards
TAG=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --prime-steps 2 --cpu-seconds 0.05 --e2e-steps 1 --no-check --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_raw.csv $B > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_raw.csv > gpurun_out/${TAG}_launches.txt; cat gpurun_out/${TAG}_launches.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_mgpath_launches_raw.csv $B --mg-path > gpurun_out/${TAG}_mgpath_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_mgpath_launches_raw.csv > gpurun_out/${TAG}_mgpath_launches.txt; cat gpurun_out/${TAG}_mgpath_launches.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tile|k_wide|k_ffat|k_slots' -c 120 --csv --log-file gpurun_out/${TAG}_facade_launches_raw.csv windflow_b200/apps/pipeline_bench.bin 128 512 > gpurun_out/${TAG}_facade_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_facade_launches_raw.csv > gpurun_out/${TAG}_facade_launches.txt; cat gpurun_out/${TAG}_facade_launches.txt
for k in k_tile_pass k_wide_scatter_ranked k_ffat_update_buckets; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 6 -c 1 -f -o gpurun_out/${TAG}_$k $B > gpurun_out/${TAG}_$k.log 2>&1
  echo "$k rc=$?"
done
