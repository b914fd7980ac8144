#!/bin/bash
# usage: prof_kernel.sh <kernel-regex> <tag> [env assignments...]  -> gpurun_out/<tag>.ncu-rep (+ raw csv)
K=$1; TAG=$2; shift 2
mkdir -p gpurun_out
env "$@" timeout 600 ncu --set full --import-source on --clock-control none -k regex:$K -s ${SKIP:-6} -c 1 -f -o gpurun_out/$TAG python bench.py --steps 2 --warmup 1 --prime-steps 1 --cpu-seconds 0.05 --e2e-steps 1 --no-check --no-extras > gpurun_out/$TAG.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/$TAG.log
