#!/bin/bash
# N GPUs: ffat tests on GPU 0, N=1 bench, then the N-GPU bench with the device timeline of the step (WFB_MG_TRACE), then the checked run
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_ffat.py tests/test_gpu_keyed.py -m gpu -q -x 2>&1 | grep -E "Error|error|assert|passed|failed" | head -12
timeout 600 python bench.py --steps 65 --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 --no-extras > gpurun_out/n1.json 2> gpurun_out/n1.err
python -c "
import json
d=json.load(open('gpurun_out/n1.json')); print('N=1', round(d['value']/1e9,2),'GT/s ms/step', round(d['ms_per_step'],4), [round(x['avg_us'],1) for x in d['roofline']['kernels']], 'check', d['check'] and d['check']['windows_compared'])" || tail -3 gpurun_out/n1.err
env WFB_MG_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 130 --warmup 3 --cpu-seconds 0.2 --e2e-steps 2 --no-extras --no-check > gpurun_out/mg_${N}_trace.json 2> gpurun_out/mg_${N}_trace.err
grep "wfb_mg rank" gpurun_out/mg_${N}_trace.err | tail -2
bash tools/mg_r2.sh $N only 2>&1 | cut -c1-200
