#!/usr/bin/env python
"""bench.py -- tuples/sec of the Map_GPU -> Filter_GPU -> Ffat_Windows_GPU (count-based) pipeline on N B200s.

    python bench.py --gpus N --steps K --warmup W                 (our arm; one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path on the host cores)

Workload (BASELINE.json north_star / SURVEY.md 8d): 64-byte tuples of the seeded synthetic stream, batch 65536,
65536 uniform keys, map (ivalue += 2, fvalue *= 1.0000001) -> filter ((ivalue & 1) == 0) -> count-based sliding
windows win 4096 / slide 64, lift {isum, fsum}, comb +. A *step* is one stream segment of `--batches-per-step`
consecutive batches handed to the operator in one call (the operator coalesces queued batches; one launch sequence
per segment). Input segments are resident in HBM in a ring larger than L2; the window state is primed (untimed) so
that every timed step is steady state (each key fires one window per 64 surviving tuples).

One JSON line is printed by rank 0 (see the contract in the task statement): value = whole-job tuples/s with
inputs resident in HBM, e2e = the same through the public call with HOST (pinned) buffers, host<->device copies
inside the timed region, roofline = dominant kernel against the measured HBM peak, cpu_baseline = the reference's
CPU path (oracle port / reference FlatFAT) timed on this box's cores on a bounded steady-state sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 65536
TUPLE_BYTES = 64
NKEYS = 65536
WIN, SLIDE = 4096, 64
MAP = dict(map_kind=1, iadd=2, fscale=1.0000001)
FILT = dict(filt_kind=1, mod=1)
SIGMA = 0.5  # selectivity of (ivalue & 1) == 0 on the synthetic stream

# algorithmic bytes. SURVEY 8d pipeline figure and the dominant kernel's own compulsory traffic.
METRIC = "tuples/sec, Map_GPU->Filter_GPU->Ffat_Windows_GPU (CB win 4096 slide 64) pipeline"
PIPELINE_BYTES_PER_TUPLE = 123.3        # read I + sigma*(3R + (O+12R)/S), I=72 R=32 O=40 S=64
INGEST_BYTES_PER_TUPLE = 64 + SIGMA * (32 + 4)   # k_tile_pass<INGEST>: read tuple, write sigma*(lifted result + slot)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML DURING the timed region (a 2 ms poll in a thread)."""

    def __init__(self, index=0):
        self.index, self.samples, self.reasons, self.max_mhz, self.ok = index, [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _poll(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.ok:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self._stop.set()
        self.t.join(timeout=1)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU Map -> Filter -> Ffat_Windows path on the host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_pipeline(kind, threads, target_seconds, keys_per_thread=64):
    """Bounded steady-state sample: `threads` replicas, each a keyby shard owning `keys_per_thread` keys and fed its own
    (already routed) tuple64 stream; state primed until every key fires windows, then timed. Returns (tuples/s, desc)."""
    from oracle import oracle as O
    n_buf = 1 << 18
    bufs = [O.gen_tuple64(s * n_buf, n_buf, O.KEY_UNIFORM, keys_per_thread) for s in range(threads)]
    pipes = [O.CpuPipe(kind, 1, 2, 1.0000001, 1, 1, WIN, SLIDE, 0, 1) for _ in range(threads)]

    def run_all(reps):
        def work(p, buf):
            for _ in range(reps):
                p.run(buf[0], buf[1], BATCH)
        th = [threading.Thread(target=work, args=(p, b)) for p, b in zip(pipes, bufs)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    prime_reps = int(np.ceil(WIN * keys_per_thread / SIGMA / n_buf)) + 1   # every key past its first window
    run_all(prime_reps)
    dt1 = run_all(1)
    reps = max(1, int(target_seconds / max(dt1, 1e-3)))
    w0 = sum(p.windows for p in pipes)
    dt = run_all(reps)
    nwin = sum(p.windows for p in pipes) - w0
    for p in pipes:
        p.close()
    tps = reps * n_buf * threads / dt
    desc = (f"{threads} replica threads x {reps} x {n_buf} tuple64 (each thread = one keyby shard with {keys_per_thread} uniform keys, "
            f"already routed; win {WIN} slide {SLIDE}; state primed to steady state; {nwin} windows in the timed sample); "
            f"{'reference wf/flatfat.hpp under the restated FFAT_Replica loop' if kind == 'reference' else 'oracle port of map.hpp/filter.hpp/ffat_replica.hpp/flatfat.hpp'}")
    return tps, desc


def best_cpu_kind():
    from oracle import oracle as O
    return "reference" if O.ref_cpu_lib() is not None else "port"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = best_cpu_kind()
    threads = min(os.cpu_count() or 1, 128)
    per_step = 12.0 / max(1, args.steps + args.warmup)
    vals = []
    desc = ""
    for i in range(args.warmup + args.steps):
        tps, desc = cpu_pipeline(kind, threads, max(1.0, per_step))
        if i >= args.warmup:
            vals.append(tps)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v,
        "unit": "tuples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (1 << 18) * threads / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i64+f64", "data": "synthetic",
        "config": {"workload": "map_filter_ffat_cb", "batch": BATCH, "tuple_bytes": TUPLE_BYTES, "keys": NKEYS, "key_dist": "uniform",
                   "win": WIN, "slide": SLIDE, "wins_per_batch": args.nb, "map": "ivalue+=2,fvalue*=1.0000001", "filter": "(ivalue&1)==0",
                   "selectivity": SIGMA,
                   "note": "the reference's CPU Map->Filter->Ffat_Windows path on this box's host cores (all of them); every step is a "
                           "bounded steady-state sample of the same stream"},
        "cpu_baseline": {"value": v, "unit": "tuples/s", "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": "tuples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from windflow_b200 import build, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- windflow_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    build.build()

    bps = args.batches_per_step
    seg_tuples = bps * BATCH
    ring = max(2, args.ring)
    nb = args.nb
    B = (nb - 1) * SLIDE + WIN
    f = ops.functors(**MAP, **FILT)
    pipelined = args.pipeline

    # ---- input ring, resident in HBM (larger than L2: ring * seg_tuples * 64 B) ---------------------------------
    # N = 1: one fused call per segment. N > 1: rank r owns the K batches [r*K, (r+1)*K) of every
    # global step, Map->Filter, partition by key % N, NCCL all-to-all, windows on the rank's key shard.
    from windflow_b200 import multigpu
    segs_whole, segs = [], []
    for r in range(ring):
        start = multigpu.owner_span(r, rank, world, seg_tuples)[0]
        b = ops.gen_tuple64(start, seg_tuples, ops.KEY_UNIFORM, NKEYS)
        b.watermark = start
        segs_whole.append(b)
        segs.append(ops.Segment([ops.DeviceBatch(b.tuples[i * BATCH * 64:(i + 1) * BATCH * 64], b.ts[i * BATCH:(i + 1) * BATCH], BATCH,
                                                 watermark=start + i * BATCH) for i in range(bps)]))
    torch.cuda.synchronize()

    if world == 1:
        ff = ops.FfatWindowsGPU(ops.PROG_TUPLE64, WIN, SLIDE, nb, max_keys=NKEYS, dense_keys=True, pipelined=pipelined)
        pipe = None
    else:
        pipe = multigpu.KeyShardedPipeline(ops, f, WIN, SLIDE, nb, NKEYS, rank, world, dev, pipelined=not args.sync_exchange)
        ff = pipe.ff
    cap = ff.max_results(seg_tuples * (2 if world > 1 else 1))
    out = torch.empty(cap * 32, dtype=torch.uint8, device=dev)
    out_ts = torch.empty(cap, dtype=torch.int64, device=dev)
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)

    def process_device_segment(whole, batches):
        if pipe is None:
            ff.process(batches, pre=f, out=out, out_ts=out_ts, n_out=n_out)
        else:
            pipe.step(batches, whole.watermark, out, out_ts, n_out)

    def step(i):
        process_device_segment(segs_whole[i % ring], segs[i % ring])

    def launches_now():
        return ff.launches + (pipe.eng.launches if pipe is not None else 0)

    # ---- prime the window state (untimed setup): every key past its first trigger --------------------------------
    prime = int(np.ceil(B * NKEYS / SIGMA / (seg_tuples * world))) + 2
    if args.prime_steps >= 0:
        prime = args.prime_steps  # profiling runs only: the timed steps are then NOT steady state
    for i in range(prime):
        step(i)
    torch.cuda.synchronize()
    it = prime
    for _ in range(args.warmup):
        step(it); it += 1
    torch.cuda.synchronize()
    windows_per_step = int(n_out.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: K steps, device-resident inputs ------------------------------------------------------------
    ff.timing(True)
    launches0 = launches_now()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(it); it += 1
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = launches_now() - launches0
    ing_ms, sort_ms, upd_ms, tot_ms, calls = ff.timing(False)
    err = ff.stats()[1]
    if err:
        raise SystemExit(f"bench.py: device error flags {err}")
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * args.steps * seg_tuples / (ms_max * 1e-3)

    # ---- e2e: the same call with HOST (pinned) buffers, copies inside the timed region ------------------------------
    e2e = run_e2e(torch, ops, process_device_segment, segs_whole, seg_tuples, bps, dev, args, world, out, n_out)

    if pipe is not None:
        pipe.flush(out, out_ts, n_out)
        torch.cuda.synchronize()
    if rank == 0:
        peak, peak_src = measured_peaks()
        ingest_ms_avg = ing_ms / max(1, calls)
        achieved = INGEST_BYTES_PER_TUPLE * seg_tuples / (ingest_ms_avg * 1e-3) / 1e9
        traffic = ncu_traffic()
        cpu_kind = best_cpu_kind()
        cpu_threads = min(os.cpu_count() or 1, 32)
        cpu_tps, cpu_desc = cpu_pipeline(cpu_kind, cpu_threads, args.cpu_seconds)
        line = {
            "metric": METRIC,
            "value": value, "unit": "tuples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i64+f64", "data": "synthetic",
            "config": {"workload": "map_filter_ffat_cb", "batch": BATCH, "tuple_bytes": TUPLE_BYTES,
                       "batches_per_step": bps, "tuples_per_step_per_gpu": seg_tuples, "keys": NKEYS, "keys_per_gpu": NKEYS // world,
                       "key_dist": "uniform", "win": WIN, "slide": SLIDE, "wins_per_batch": nb,
                       "map": "ivalue+=2,fvalue*=1.0000001", "filter": "(ivalue&1)==0", "selectivity": SIGMA,
                       "l2": f"inputs larger than L2: ring of {ring} segments x {seg_tuples * 64 / 1e6:.0f} MB",
                       "state_primed_steps": prime, "windows_per_step_per_gpu": windows_per_step,
                       "pipelined": args.pipeline if world == 1 else (not args.sync_exchange),
                       "parallelism": f"keyby{world}" + ("" if world == 1 else " (Map->Filter->lift + partition by key % N | NCCL all-to-all of 32-B results | Ffat on the key shard, records read in place)")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": e2e,
            "roofline": {"bound": "hbm",
                         "kernel": "k_tile_pass<ProgTuple64, MODE_INGEST>" if world == 1 else
                                   "whole pipeline per GPU, SURVEY 8d bytes (the kernel-level roofline is the N=1 line: at N>1 the timed handle is the destination side)",
                         "achieved": achieved if world == 1 else value / world * PIPELINE_BYTES_PER_TUPLE / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": (achieved if world == 1 else value / world * PIPELINE_BYTES_PER_TUPLE / 1e9) / peak,
                         "traffic": (traffic or {}).get("ingest_dram_bytes_per_launch") if world == 1 else None,
                         "peak_source": peak_src, "bytes_per_tuple": INGEST_BYTES_PER_TUPLE,
                         "avg_launch_ms": ingest_ms_avg,
                         "phase_ms_per_step": {"ingest": ing_ms / max(1, calls), "offsets+sort": sort_ms / max(1, calls),
                                               "update": upd_ms / max(1, calls), "call": tot_ms / max(1, calls)},
                         "pipeline": {"bytes_per_tuple": PIPELINE_BYTES_PER_TUPLE,
                                      "achieved": value / world * PIPELINE_BYTES_PER_TUPLE / 1e9,
                                      "frac": value / world * PIPELINE_BYTES_PER_TUPLE / 1e9 / peak}},
            "cpu_baseline": {"value": cpu_tps, "unit": "tuples/s", "cores": cpu_threads, "kind": cpu_kind, "sample": cpu_desc},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(torch, ops, process_device_segment, segs_whole, seg_tuples, bps, dev, args, world, out, n_out):
    """Same operator call(s), inputs start in pinned host memory every step; result count + results come back."""
    import torch.distributed as dist
    steps = max(2, min(args.steps, args.e2e_steps))
    nbuf = 2
    host_t = [torch.empty(seg_tuples * 64, dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
    host_ts = [torch.empty(seg_tuples, dtype=torch.int64).pin_memory() for _ in range(nbuf)]
    for k in range(nbuf):  # the segments' bytes are copied out to the host once, untimed
        host_t[k].copy_(segs_whole[k % len(segs_whole)].tuples); host_ts[k].copy_(segs_whole[k % len(segs_whole)].ts)
    dev_t = [torch.empty(seg_tuples * 64, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    dev_ts = [torch.empty(seg_tuples, dtype=torch.int64, device=dev) for _ in range(nbuf)]
    host_n = torch.zeros(1, dtype=torch.int32).pin_memory()
    res_cap = out.numel() // 32
    host_res = torch.empty(res_cap * 32, dtype=torch.uint8).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    ready = [torch.cuda.Event() for _ in range(nbuf)]
    freed = [torch.cuda.Event() for _ in range(nbuf)]
    wm0 = segs_whole[0].watermark

    def h2d(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[k])
            dev_t[k].copy_(host_t[k], non_blocking=True)
            dev_ts[k].copy_(host_ts[k], non_blocking=True)
            ready[k].record(copy_stream)

    def compute(k, step_idx):
        main.wait_event(ready[k])
        wm = wm0 + step_idx * seg_tuples
        whole = ops.DeviceBatch(dev_t[k], dev_ts[k], seg_tuples, wm)
        batches = [ops.DeviceBatch(dev_t[k][i * BATCH * 64:(i + 1) * BATCH * 64], dev_ts[k][i * BATCH:(i + 1) * BATCH], BATCH,
                                   watermark=wm + i * BATCH) for i in range(bps)]
        process_device_segment(whole, batches)
        freed[k].record(main)
        host_n.copy_(n_out, non_blocking=True)

    d2h_bytes = 0
    for k in range(nbuf):
        freed[k].record(main)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h2d(0)
    for s in range(steps):
        k = s % nbuf
        if s + 1 < steps:
            h2d((s + 1) % nbuf)
        compute(k, s)
        main.synchronize()                       # the caller reads the step's result count ...
        nres = int(host_n.item())
        if nres:                                  # ... and the window results themselves
            host_res[:nres * 32].copy_(out[:nres * 32], non_blocking=True)
            d2h_bytes += nres * 32
        d2h_bytes += 4
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    val = world * steps * seg_tuples / (float(t_ms.item()) * 1e-3)
    return {"value": val, "unit": "tuples/s", "h2d_bytes_per_step": seg_tuples * 72, "d2h_bytes_per_step": d2h_bytes // steps,
            "steps": steps, "note": "pinned host segment -> H2D (double-buffered on a copy stream) -> the operator call(s) "
                                    "-> D2H of the result count and the window results"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batches-per-step", type=int, default=128, help="queued batches the replica hands to the operator per call (one stream segment)")
    ap.add_argument("--ring", type=int, default=4)
    ap.add_argument("--nb", type=int, default=65, help="withNumWinPerBatch")
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--pipeline", action="store_true", help="WFB_FFAT_PIPELINED handle: results one call late, sort+update overlap the next ingest")
    ap.add_argument("--sync-exchange", action="store_true", help="N > 1: exchange and window update of a step right after its source pass (no overlap with the next step)")
    ap.add_argument("--prime-steps", type=int, default=-1, help="override state priming (ncu runs); default: steady state")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
