#!/bin/bash
mkdir -p gpurun_out
R() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 40 --warmup 5 --cpu-seconds 0.2 --e2e-steps 2 --no-extras --no-check --ring 60 > gpurun_out/host_$tag.json 2> gpurun_out/host_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/host_$tag.json')); print('$tag', d['ms_per_step'], d['host_issue_ms_per_step'], d['roofline']['avg_launch_ms'])"; }
R base
R noadd WFB_BENCH_NOADD=1
R notiming WFB_BENCH_NOADD=1 WFB_BENCH_NOTIMING=1
