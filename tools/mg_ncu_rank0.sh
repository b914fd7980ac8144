#!/bin/bash
# launched by torchrun as the per-rank program: rank 0 runs under ncu (launch list: per-kernel durations at N>1), the others plainly
if [ "$LOCAL_RANK" == "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/mgN_rank0_launches_raw.csv python "$@"
else
  exec python "$@"
fi
