#!/usr/bin/env python
"""Per-phase device times of the multi-GPU pipeline object (torchrun, any world size): source pass, exchange, update."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from windflow_b200 import ops, multigpu
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
BATCH, NKEYS, WIN, SLIDE, NB, BPS = 65536, 65536, 4096, 64, 65, 64
f = ops.functors(map_kind=1, iadd=2, fscale=1.0000001, filt_kind=1, mod=1)
pipe = multigpu.KeyShardedPipeline(ops, f, WIN, SLIDE, NB, NKEYS, rank, world, dev, pipelined=False)
cap = pipe.ff.max_results(BPS * BATCH * 2)
out = torch.empty(cap * 32, dtype=torch.uint8, device=dev); out_ts = torch.empty(cap, dtype=torch.int64, device=dev)
n_out = torch.zeros(1, dtype=torch.int32, device=dev)
segs = []
for r in range(3):
    start = multigpu.owner_span(r, rank, world, BPS * BATCH)[0]
    b = ops.gen_tuple64(start, BPS * BATCH, ops.KEY_UNIFORM, NKEYS)
    segs.append([ops.DeviceBatch(b.tuples[i * BATCH * 64:(i + 1) * BATCH * 64], b.ts[i * BATCH:(i + 1) * BATCH], BATCH, watermark=start + i * BATCH) for i in range(BPS)])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
acc = np.zeros(3); n = 0
for step in range(40):
    sl = pipe.slots[pipe.step_no & 1]; pipe.step_no += 1
    main = torch.cuda.current_stream(dev)
    ev[0].record(main)
    pipe._source(sl, segs[step % 3], step)
    main.wait_stream(pipe.comm); ev[1].record(main)
    ex = pipe._exchange(sl)
    main.wait_event(sl.ev_a2a); ev[2].record(main)
    pipe._update(sl, *ex, out, out_ts, n_out)
    ev[3].record(main)
    torch.cuda.synchronize()
    if step >= 25:
        acc += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]; n += 1
if rank == 0:
    print("world %d  source %.3f ms  exchange %.3f ms  update %.3f ms" % (world, *(acc / n)))
    pipe.ff.timing(True)
dist.barrier()
