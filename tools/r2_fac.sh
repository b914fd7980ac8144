#!/bin/bash
# facade pipeline: timing at K=128 / 64 / 16 and the launch list at K=128; then the N>1 step with one rank
mkdir -p gpurun_out
for K in 128 64 16; do windflow_b200/apps/pipeline_bench.bin $K 1024 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('facade K=$K', round(d['tuples_per_s']/1e9,2), 'GT/s')"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tile|k_wide|k_ffat|k_slots' -c 120 --csv --log-file gpurun_out/facade_launches_raw.csv windflow_b200/apps/pipeline_bench.bin 128 512 > gpurun_out/facade_launches.log 2>&1
python tools/launch_summary.py gpurun_out/facade_launches_raw.csv > gpurun_out/facade_launches.txt; cat gpurun_out/facade_launches.txt
R() { tag=$1; shift
timeout 600 python bench.py "$@" --steps 65 --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 --no-extras > gpurun_out/mg1_$tag.json 2> gpurun_out/mg1_$tag.err
python -c "
import json
for l in open('gpurun_out/mg1_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],4), 'host', round(d.get('host_issue_ms_per_step'),3), [round(x['avg_us'],1) for x in (d['roofline'].get('kernels') or [])], 'check', d['check'] and d['check']['windows_compared'])
" || tail -5 gpurun_out/mg1_$tag.err
}
R mgpath --mg-path
R default
