#!/bin/bash
# one-GPU look at the N>1 step (bench.py --mg-path, one rank): tests, timing of the bucketed and the two-partition exchange, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_keyed.py tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/mg1_pytest.log 2>&1; echo "pytest rc=$? $(grep -E 'passed|failed' gpurun_out/mg1_pytest.log | tail -1)"
R() { tag=$1; shift
env $ENVV timeout 600 python bench.py "$@" --steps 65 --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 --no-extras > gpurun_out/mg1_$tag.json 2> gpurun_out/mg1_$tag.err
python -c "
import json
for l in open('gpurun_out/mg1_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],4), 'host', round(d.get('host_issue_ms_per_step'),3), [round(x['avg_us'],1) for x in (d['roofline'].get('kernels') or [])], 'check', d['check'] and d['check']['windows_compared'])
" || tail -5 gpurun_out/mg1_$tag.err
}
R mgpath --mg-path
ENVV="WFB_MG_BUCKETED=0" R mgpath_twopart --mg-path
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/mg1_launches_raw.csv python bench.py --mg-path --steps 2 --warmup 3 --prime-steps 2 --cpu-seconds 0.05 --e2e-steps 2 --no-check --no-extras > gpurun_out/mg1_launches.log 2>&1
python tools/launch_summary.py gpurun_out/mg1_launches_raw.csv > gpurun_out/mg1_launches.txt; cat gpurun_out/mg1_launches.txt
