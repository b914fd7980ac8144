#!/bin/bash
# ncu launch list (per-launch durations, cold cache, serialised) of a short bench run: gpurun_out/<tag>_launches.txt
TAG=${1:-r2}; shift
mkdir -p gpurun_out
env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_raw.csv python bench.py --steps 2 --warmup 3 --prime-steps 2 --cpu-seconds 0.05 --e2e-steps 2 --no-check --no-extras > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_raw.csv > gpurun_out/${TAG}_launches.txt; cat gpurun_out/${TAG}_launches.txt
