import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
from windflow_b200 import ops, multigpu, build
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
BATCH=65536; bps=64; seg=bps*BATCH
f=ops.functors(map_kind=1,iadd=2,fscale=1.0000001,filt_kind=1)
b=ops.gen_tuple64(rank*seg, seg, ops.KEY_UNIFORM, 65536)
batches=[ops.DeviceBatch(b.tuples[i*BATCH*64:(i+1)*BATCH*64], None, BATCH, i) for i in range(bps)]
pipe=multigpu.KeyShardedPipeline(ops,f,4096,64,65,65536,rank,world,dev,pipelined=True)
cap=pipe.ff.max_results(seg*2); out=torch.empty(cap*32,dtype=torch.uint8,device=dev); ots=torch.empty(cap,dtype=torch.int64,device=dev); n_out=torch.zeros(1,dtype=torch.int32,device=dev)
for _ in range(5): pipe.step(batches,0,out,ots,n_out)
torch.cuda.synchronize(); dist.barrier()
# phase timing with host clocks + syncs
T={}
def tick(name,t0):
    torch.cuda.synchronize(); T[name]=T.get(name,0)+time.perf_counter()-t0
N=20
for _ in range(N):
    t0=time.perf_counter(); pipe._ensure(seg); pipe.eng.shard_lift(batches,f,world,pipe.regions,pipe.region_cap,pipe.counts); tick("shard_lift",t0)
    t0=time.perf_counter(); cnt=pipe.counts.cpu().tolist(); tick("counts_d2h",t0)
    t0=time.perf_counter(); sc=torch.tensor(cnt[:world],dtype=torch.int64,device=dev); rc,rw=multigpu.exchange_counts(sc,0); rch=rc.cpu().tolist(); tick("exch_counts",t0)
    t0=time.perf_counter(); pipe.recv,offs=multigpu.exchange_regions(pipe.regions,pipe.region_cap,cnt[:world],rch,32,pipe.recv); tick("a2a",t0)
    t0=time.perf_counter(); chunks=[ops.DeviceBatch(pipe.recv[offs[s]*32:offs[s+1]*32],None,offs[s+1]-offs[s],0) for s in range(world)]; pipe.ff.process(chunks,pre=None,out=out,out_ts=ots,n_out=n_out); tick("ffat",t0)
if rank==0: print({k:round(v/N*1e3,3) for k,v in T.items()}, "ms per step; recv", sum(rch))
dist.destroy_process_group()
