#!/usr/bin/env python
"""Per-config numbers of SURVEY.md section 8(d) next to the headline pipeline of bench.py (one JSON line per config):
cfg 2  Map_GPU -> Filter_GPU fused, batch 65536 x 64-byte tuples            (algorithmic 72*(1+sigma) = 108 B/tuple)
cfg 3  Reduce_GPU keyed, 1M keys Zipf-0.8, batch 65536                      (algorithmic 72*(1+d) B/tuple, d measured)
cfg 4  Ffat_Windows_GPU CB win 4096 slide 64, 65536 keys, Nb = 65 and Nb = 1 (algorithmic 174.6 B/tuple)
Inputs are a ring of device-resident batches larger than L2; CUDA events over the timed calls; no oracle involved."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from windflow_b200 import ops

BATCH = 65536


def zipf_cdf(nkeys, s=0.8):  # same table as the oracle's generator (oracle/oracle.py:85) without importing it
    w = 1.0 / np.power(np.arange(1, nkeys + 1, dtype=np.float64), s)
    c = np.cumsum(w); c /= c[-1]
    return c


def timed(fn, iters, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(warm + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=200); a = ap.parse_args()
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6569.6) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6569.6
    dev = torch.device("cuda", 0)
    f = ops.functors(map_kind=1, iadd=2, fscale=1.0000001, filt_kind=1, mod=1)
    ring = 64  # 64 batches x 4.7 MB = 300 MB > L2
    # ---- cfg 2 --------------------------------------------------------------------------------------------------
    eng = ops.Engine(ops.PROG_TUPLE64)
    batches = [ops.gen_tuple64(i * BATCH, BATCH, ops.KEY_RR, 65536) for i in range(ring)]
    out = ops.DeviceBatch(torch.empty_like(batches[0].tuples), torch.empty_like(batches[0].ts), BATCH, 0)
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)
    ms = timed(lambda i: eng.map_filter(batches[i % ring], f, out=out, n_out=n_out), a.iters)
    sigma = int(n_out.item()) / BATCH
    tps = BATCH / (ms * 1e-3); bpt = 72 * (1 + sigma)
    print(json.dumps({"config": "cfg2 Map_GPU->Filter_GPU fused, one call per batch of 65536 x 64 B", "tuples_per_s": tps, "ms_per_batch": ms,
                      "selectivity": sigma, "bytes_per_tuple": bpt, "achieved_gbs": tps * bpt / 1e9, "frac_of_measured_peak": tps * bpt / 1e9 / peak}))
    ob = [ops.DeviceBatch(torch.empty_like(b.tuples), torch.empty_like(b.ts), BATCH, 0) for b in batches]
    nos = torch.zeros(ring, dtype=torch.int32, device=dev)
    seg_in, seg_out = ops.Segment(batches), ops.Segment(ob)
    ms = timed(lambda i: eng.map_filter_batches(seg_in, f, seg_out, nos), max(10, a.iters // 4))
    tps = ring * BATCH / (ms * 1e-3)
    print(json.dumps({"config": f"cfg2 Map_GPU->Filter_GPU fused, {ring} queued batches of 65536 x 64 B per call (wfb_map_filter_batches)", "tuples_per_s": tps,
                      "ms_per_call": ms, "selectivity": sigma, "bytes_per_tuple": bpt, "achieved_gbs": tps * bpt / 1e9, "frac_of_measured_peak": tps * bpt / 1e9 / peak}))
    # ---- cfg 3 --------------------------------------------------------------------------------------------------
    cdf = torch.from_numpy(zipf_cdf(1000000)).to(dev)
    zb = [ops.gen_tuple64(i * BATCH, BATCH, ops.KEY_ZIPF, 1000000, zipf_cdf=cdf) for i in range(ring)]
    eng3 = ops.Engine(ops.PROG_TUPLE64); eng3.set_key_bits(20)
    ms = timed(lambda i: eng3.reduce_by_key(zb[i % ring], out=out, n_out=n_out), a.iters)
    d = int(n_out.item()) / BATCH
    tps = BATCH / (ms * 1e-3); bpt = 72 * (1 + d)
    print(json.dumps({"config": "cfg3 Reduce_GPU keyed, 1M keys Zipf-0.8, one call per batch of 65536", "tuples_per_s": tps, "ms_per_batch": ms,
                      "distinct_fraction": d, "bytes_per_tuple": bpt, "achieved_gbs": tps * bpt / 1e9, "frac_of_measured_peak": tps * bpt / 1e9 / peak}))
    zo = [ops.DeviceBatch(torch.empty_like(b.tuples), torch.empty_like(b.ts), BATCH, 0) for b in zb]
    zin, zout = ops.Segment(zb), ops.Segment(zo)
    ms = timed(lambda i: eng3.reduce_by_key_batches(zin, zout, nos), max(10, a.iters // 8))
    tps = ring * BATCH / (ms * 1e-3)
    print(json.dumps({"config": f"cfg3 Reduce_GPU keyed, 1M keys Zipf-0.8, {ring} queued batches per call (wfb_reduce_by_key_batches)", "tuples_per_s": tps,
                      "ms_per_call": ms, "distinct_fraction": float(nos.sum().item()) / (ring * BATCH), "bytes_per_tuple": bpt,
                      "achieved_gbs": tps * bpt / 1e9, "frac_of_measured_peak": tps * bpt / 1e9 / peak}))
    # ---- keyed-stateful Map_GPU / Filter_GPU: 65536 uniform keys, 64 queued batches per call --------------------
    kb = [ops.gen_tuple64(i * BATCH, BATCH, ops.KEY_UNIFORM, 65536) for i in range(ring)]
    kseg = ops.Segment(kb)
    ksm = ops.KeyedState(ops.PROG_TUPLE64, max_keys=65536, dense_keys=True)
    ms = timed(lambda i: ksm.map(kseg, f), max(10, a.iters // 8))
    tps = ring * BATCH / (ms * 1e-3)
    print(json.dumps({"config": f"Map_GPU keyed-stateful (counter per key), 65536 uniform keys, {ring} queued batches per call (wfb_map_stateful)",
                      "tuples_per_s": tps, "ms_per_call": ms, "bytes_per_tuple": 128, "achieved_gbs": tps * 128 / 1e9, "frac_of_measured_peak": tps * 128 / 1e9 / peak}))
    ksf = ops.KeyedState(ops.PROG_TUPLE64, max_keys=65536, dense_keys=True)
    ms = timed(lambda i: ksf.filter(kseg, ops.functors(filt_kind=1), seg_out, nos), max(10, a.iters // 8))
    tps = ring * BATCH / (ms * 1e-3)
    print(json.dumps({"config": f"Filter_GPU keyed-stateful, 65536 uniform keys, {ring} queued batches per call (wfb_filter_stateful)",
                      "tuples_per_s": tps, "ms_per_call": ms, "selectivity": float(nos.sum().item()) / (ring * BATCH)}))
    # ---- time-based windows, one batch per call: ts = tuple index, keys round-robin, so a key sees one tuple every
    # nk timestamp units; win = 4096 nk, slide = 64 nk are the count-based config's windows (4096 / 64 tuples per key) in time units
    for nk in (65536, 1024):
        tbh = ops.FfatWindowsGPU(ops.PROG_TUPLE64, 4096 * nk, 64 * nk, 65, max_keys=nk, dense_keys=True, win_type=1)
        tbb = [ops.gen_tuple64(i * BATCH, BATCH, ops.KEY_RR, nk) for i in range(ring)]
        for i, b in enumerate(tbb):
            b.watermark = i * BATCH
        cap = 1 << 22
        o = torch.empty(cap * 32, dtype=torch.uint8, device=dev); ots = torch.empty(cap, dtype=torch.int64, device=dev)
        state = {"i": 0}
        def tb_step(_):
            # a fresh stretch of the stream every call (timestamps keep growing): regenerate into the ring slot
            i = state["i"]; state["i"] += 1
            b = tbb[i % ring]
            ops.gen_tuple64(i * BATCH, BATCH, ops.KEY_RR, nk, tuples=b.tuples, ts=b.ts)
            b.watermark = i * BATCH
            tbh.process([b], out=o, out_ts=ots, n_out=n_out)
        ms = timed(tb_step, max(20, a.iters // 4))
        gen_ms = timed(lambda i: ops.gen_tuple64(i * BATCH, BATCH, ops.KEY_RR, nk, tuples=tbb[0].tuples, ts=tbb[0].ts), 50)
        tps = BATCH / ((ms - gen_ms) * 1e-3)
        print(json.dumps({"config": f"Ffat_Windows_GPU time-based, win 4096*{nk} slide 64*{nk} ts units (= 4096 / 64 tuples per key), Nb=65, {nk} round-robin keys, one batch of 65536 per call (first version, stream syncs per batch)",
                          "tuples_per_s": tps, "ms_per_batch": ms - gen_ms, "windows_last_call": int(n_out.item())}))
        assert tbh.stats()[1] == 0
        tbh.close()
    # ---- cfg 4 --------------------------------------------------------------------------------------------------
    for nb in (65, 1):
        for bps in (1, 64):
            ff = ops.FfatWindowsGPU(ops.PROG_TUPLE64, 4096, 64, nb, max_keys=65536, dense_keys=True)
            nseg = 4 if bps == 64 else 64
            segs = []
            for r in range(nseg):
                b = ops.gen_tuple64(r * bps * BATCH, bps * BATCH, ops.KEY_UNIFORM, 65536)
                segs.append(ops.Segment([ops.DeviceBatch(b.tuples[i * BATCH * 64:(i + 1) * BATCH * 64], b.ts[i * BATCH:(i + 1) * BATCH], BATCH, watermark=i) for i in range(bps)]))
            cap = ff.max_results(bps * BATCH)
            o = torch.empty(cap * 32, dtype=torch.uint8, device=dev); ots = torch.empty(cap, dtype=torch.int64, device=dev)
            prime = int(np.ceil(((nb - 1) * 64 + 4096) * 65536 / (bps * BATCH))) + 2  # every key past its first trigger
            for i in range(prime):
                ff.process(segs[i % nseg], out=o, out_ts=ots, n_out=n_out)
            ms = timed(lambda i: ff.process(segs[i % nseg], out=o, out_ts=ots, n_out=n_out), max(10, a.iters // (4 if bps == 64 else 1)))
            tps = bps * BATCH / (ms * 1e-3); bpt = 174.6
            print(json.dumps({"config": f"cfg4 Ffat_Windows_GPU CB win 4096 slide 64, 65536 uniform keys, Nb={nb}, {bps} batch(es) per call", "tuples_per_s": tps,
                              "ms_per_call": ms, "windows_per_call": int(n_out.item()), "bytes_per_tuple": bpt, "achieved_gbs": tps * bpt / 1e9,
                              "frac_of_measured_peak": tps * bpt / 1e9 / peak}))
            assert ff.stats()[1] == 0
            ff.close()


if __name__ == "__main__":
    main()
