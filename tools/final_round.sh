#!/bin/bash
# the round's closing run on one GPU: full GPU suite, smoke, default bench line (+ reference arm), profile artefacts, per-config numbers
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? $(grep -E 'passed|failed' gpurun_out/${TAG}_pytest_gpu.log | tail -1)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(round(d['value']/1e9,2),'GT/s', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']/1e9,3), 'cpu', d['cpu_baseline']['value'], 'check', d['check'] and d['check']['passed'])" || tail -3 gpurun_out/${TAG}_bench.err
bash tools/profile_round.sh $TAG > gpurun_out/${TAG}_profile_round.log 2>&1; tail -4 gpurun_out/${TAG}_profile_round.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_reference.json
timeout 600 python tools/bench_configs.py --iters 100 > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$? lines=$(wc -l < gpurun_out/${TAG}_configs.jsonl)"
