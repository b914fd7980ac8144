"""GPU parity tests (-m gpu): Reduce_GPU (keyed / un-keyed), KeyBy_Emitter_GPU grouping and the key -> shard
partition through the C ABI against the oracle. Integer results, indices and order are bit-exact; floating-point
sums within 1e-6 relative (thrust's association is implementation-defined)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SIZES = [1, 2, 33, 255, 2048, 2049, 65536, 100001]


def _batch(O, n, nkeys, mode=None, start=0):
    t, ts = O.gen_tuple64(start, n, O.KEY_UNIFORM if mode is None else mode, nkeys)
    t["pad"] = np.arange(n * 4, dtype=np.uint64).reshape(n, 4)
    return t, ts


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("nkeys", [1, 7, 5000])
def test_keyby_group(wfb, oracle, n, nkeys):
    import torch
    O, ops = oracle, wfb
    t, ts = _batch(O, n, nkeys)
    t["key"] = t["key"] * 0x9E3779B97F4A7C15 % (1 << 40)  # scattered 40-bit keys
    eng = ops.Engine(ops.PROG_TUPLE64)
    eng.set_key_bits(40)
    start, mp, dk, nk = eng.keyby_group(ops.DeviceBatch.from_host(t, ts))
    torch.cuda.synchronize()
    es, em, ek = O.keyby_group(t["key"], 1)  # GPU->GPU path: ascending key order (keyby_emitter_gpu.hpp:547-564)
    k = int(nk.item())
    assert k == len(ek)
    assert np.array_equal(dk.cpu().numpy().view(np.uint64)[:k], ek)
    assert np.array_equal(start.cpu().numpy()[:k], es)
    assert np.array_equal(mp.cpu().numpy()[:n], em)


def test_keyby_group_full_64bit_keys(wfb, oracle):
    import torch
    O, ops = oracle, wfb
    n = 30000
    t, ts = _batch(O, n, 300)
    t["key"] = (t["key"] + 1) * np.uint64(0xD6E8FEB86659FD93)  # wraps: uses all 64 bits
    eng = ops.Engine(ops.PROG_TUPLE64)
    start, mp, dk, nk = eng.keyby_group(ops.DeviceBatch.from_host(t, ts))
    torch.cuda.synchronize()
    es, em, ek = O.keyby_group(t["key"], 1)
    k = int(nk.item())
    assert k == len(ek) and np.array_equal(dk.cpu().numpy().view(np.uint64)[:k], ek)
    assert np.array_equal(start.cpu().numpy()[:k], es) and np.array_equal(mp.cpu().numpy()[:n], em)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("dist", ["uniform50", "zipf1m", "single"])
def test_reduce_by_key(wfb, oracle, n, dist):
    import torch
    O, ops = oracle, wfb
    if dist == "uniform50":
        t, ts = _batch(O, n, 50)
    elif dist == "zipf1m":
        t, ts = _batch(O, n, 1000000, O.KEY_ZIPF)   # BASELINE config 3: 1M keys, Zipf 0.8
    else:
        t, ts = _batch(O, n, 1)
    ts = (ts * 7919) % 100003  # non-monotone timestamps: ts of the result = max over the key
    eng = ops.Engine(ops.PROG_TUPLE64)
    eng.set_key_bits(20)
    out, n_out = eng.reduce_by_key(ops.DeviceBatch.from_host(t, ts))
    torch.cuda.synchronize()
    exp, ets = O.reduce_tuple64(t, ts)
    k = int(n_out.item())
    assert k == len(exp)
    got = ops.to_host(out.tuples, ops.TUPLE64)[:k]
    assert np.array_equal(got["key"], exp["key"])            # ascending key order
    assert np.array_equal(got["ivalue"], exp["ivalue"])
    assert np.allclose(got["fvalue"], exp["fvalue"], rtol=1e-6, atol=0)
    assert np.array_equal(got["id"], exp["id"]) and np.array_equal(got["pad"], exp["pad"])  # singletons pass through
    assert np.array_equal(ops.ts_to_host(out.ts)[:k], ets)


@pytest.mark.parametrize("sizes", [[65536], [3000, 0, 1, 65536, 257, 40000, 5], [1000] * 9])
def test_reduce_by_key_batches(wfb, oracle, sizes):
    """Reduce_GPU over K queued batches in one launch sequence == the oracle per batch (ascending keys, ts = max)."""
    import torch
    O, ops = oracle, wfb
    eng = ops.Engine(ops.PROG_TUPLE64)
    eng.set_key_bits(20)
    ins, outs, hosts = [], [], []
    for i, n in enumerate(sizes):
        t, ts = O.gen_tuple64(1000 * i, n, O.KEY_ZIPF if i % 2 == 0 else O.KEY_UNIFORM, 1000000 if i % 2 == 0 else 37)
        ts = (ts * 7919) % 100003
        hosts.append((t, ts))
        if n:
            b = ops.DeviceBatch.from_host(t, ts)
        else:
            b = ops.DeviceBatch(torch.empty(0, dtype=torch.uint8, device="cuda"), torch.empty(0, dtype=torch.int64, device="cuda"), 0, 0)
        ins.append(b)
        outs.append(ops.DeviceBatch(torch.empty_like(b.tuples), torch.empty_like(b.ts), n, 0))
    n_out = torch.full((len(sizes),), 777, dtype=torch.int32, device="cuda")
    eng.reduce_by_key_batches(ins, outs, n_out)
    torch.cuda.synchronize()
    no = n_out.cpu().numpy()
    for i, (t, ts) in enumerate(hosts):
        if len(t) == 0:
            assert no[i] == 0
            continue
        exp, ets = O.reduce_tuple64(t, ts)
        assert no[i] == len(exp), (i, no[i], len(exp))
        got = ops.to_host(outs[i].tuples, ops.TUPLE64)[:len(exp)]
        assert np.array_equal(got["key"], exp["key"]) and np.array_equal(got["ivalue"], exp["ivalue"])
        assert np.allclose(got["fvalue"], exp["fvalue"], rtol=1e-6, atol=0)
        assert np.array_equal(got["id"], exp["id"]) and np.array_equal(got["pad"], exp["pad"])
        assert np.array_equal(ops.ts_to_host(outs[i].ts)[:len(exp)], ets)


@pytest.mark.parametrize("n", [1, 100, 1024, 65536, 77777])
def test_reduce_all(wfb, oracle, n):
    import torch
    O, ops = oracle, wfb
    t, ts = _batch(O, n, 100)
    eng = ops.Engine(ops.PROG_TUPLE64)
    out_t, out_ts = eng.reduce_all(ops.DeviceBatch.from_host(t, ts))
    torch.cuda.synchronize()
    r = ops.to_host(out_t, ops.TUPLE64)[0]
    # thrust::reduce(init = default item): key of the init item (0), sums over the batch, ts = max (reduce_gpu.hpp:264-273)
    assert r["key"] == 0 and r["id"] == 0
    assert r["ivalue"] == t["ivalue"].sum()
    assert abs(r["fvalue"] - t["fvalue"].sum()) <= 1e-6 * abs(t["fvalue"].sum())
    assert int(ops.ts_to_host(out_ts)[0]) == int(ts.max())


def test_reduce_reference_test_functor(wfb, oracle):
    """Reduce_Functor_GPU of tests/graph_tests_gpu/graph_common_gpu.hpp:268-279 on {key, value} tuples."""
    import torch
    O, ops = oracle, wfb
    rng = np.random.default_rng(4)
    n = 20000
    t = np.zeros(n, dtype=ops.WFTEST16)
    t["key"] = rng.integers(0, 9, n)
    t["value"] = rng.integers(-100, 100, n)
    ts = np.arange(n, dtype=np.uint64)
    eng = ops.Engine(ops.PROG_WFTEST16)
    eng.set_key_bits(8)
    out, n_out = eng.reduce_by_key(ops.DeviceBatch.from_host(t, ts))
    torch.cuda.synchronize()
    k = int(n_out.item())
    got = ops.to_host(out.tuples, ops.WFTEST16)[:k]
    uk = np.unique(t["key"])
    assert np.array_equal(got["key"], uk)
    assert np.array_equal(got["value"], [t["value"][t["key"] == x].sum() for x in uk])
    assert np.array_equal(ops.ts_to_host(out.ts)[:k], [ts[t["key"] == x].max() for x in uk])


@pytest.mark.parametrize("n", [1, 257, 65536, 100001])
@pytest.mark.parametrize("shards", [1, 2, 4, 8])
def test_shard_by_key(wfb, oracle, n, shards):
    import torch
    O, ops = oracle, wfb
    t, ts = _batch(O, n, 65536)
    eng = ops.Engine(ops.PROG_TUPLE64)
    out, seg = eng.shard_by_key(ops.DeviceBatch.from_host(t, ts), shards)
    torch.cuda.synchronize()
    dest = O.route(t["key"], shards)          # key % num_dests (keyby_emitter.hpp:215-217)
    order = np.argsort(dest, kind="stable")   # stable: arrival order inside a shard
    got = ops.to_host(out.tuples, ops.TUPLE64)
    assert got.tobytes() == t[order].tobytes()
    assert np.array_equal(ops.ts_to_host(out.ts), ts[order])
    off = seg.cpu().numpy()
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(np.bincount(dest, minlength=shards))]))


@pytest.mark.parametrize("shards", [1, 2, 3, 8])
@pytest.mark.parametrize("sizes", [[1], [255, 257, 0, 1000], [65536, 65536, 4097]])
def test_shard_lift_fused(wfb, oracle, shards, sizes):
    """Fused Map -> Filter -> lift -> stable partition by key % n == the operators applied one after the other."""
    import torch
    O, ops = oracle, wfb
    n = sum(sizes)
    t, ts = _batch(O, n, 5000)
    f = ops.functors(map_kind=1, iadd=2, fscale=1.0000001, filt_kind=1)
    eng = ops.Engine(ops.PROG_TUPLE64)
    batches, off = [], 0
    for sz in sizes:
        if sz:
            batches.append(ops.DeviceBatch.from_host(t[off:off + sz], ts[off:off + sz]))
        else:
            batches.append(ops.DeviceBatch(torch.empty(0, dtype=torch.uint8, device="cuda"), None, 0))
        off += sz
    cap = n
    regions = torch.zeros(shards * cap * 32, dtype=torch.uint8, device="cuda")
    counts = torch.zeros(9, dtype=torch.int32, device="cuda")
    eng.shard_lift(batches, f, shards, regions, cap, counts)
    torch.cuda.synchronize()
    surv, _, _ = O.map_filter_tuple64(t, ts, 1, 2, 1.0000001, 1)
    lifted = O.lift_tuple64(surv)
    dest = O.route(surv["key"], shards)
    c = counts.cpu().numpy()
    assert c[8] == 0
    got = ops.to_host(regions, ops.RESULT32).reshape(shards, cap)
    for d in range(shards):
        exp = lifted[dest == d]
        assert c[d] == len(exp)
        assert got[d][:c[d]].tobytes() == exp.tobytes()


@pytest.mark.parametrize("pipelined", [False, True])
def test_key_sharded_pipeline_world1_nccl(wfb, oracle, pipelined):
    """The multi-GPU pipeline object on a world of one rank (NCCL): shard_lift -> all-to-all -> lifted-record window
    operator (reading the received records in place) must equal the oracle windows of the fused single-GPU path."""
    import os
    import torch
    import torch.distributed as dist
    from windflow_b200 import multigpu
    O, ops = oracle, wfb
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    win, slide, nb, nkeys, n, batch = 64, 16, 2, 40, 60000, 4096
    t, ts = O.gen_tuple64(0, n, O.KEY_UNIFORM, nkeys)
    f = ops.functors(map_kind=1, iadd=2, fscale=1.0000001, filt_kind=1)
    pipe = multigpu.KeyShardedPipeline(ops, f, win, slide, nb, 64, 0, 1, torch.device("cuda", 0), pipelined=pipelined)
    go = O.FfatGpuOracle(win, slide, nb)
    cap = pipe.ff.max_results(n)
    out = torch.empty(cap * 32, dtype=torch.uint8, device="cuda"); out_ts = torch.empty(cap, dtype=torch.int64, device="cuda")
    n_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    got, exp = [], []
    step = 3 * batch
    for s0 in range(0, n, step):
        bs = [ops.DeviceBatch.from_host(t[b:b + batch], ts[b:b + batch]) for b in range(s0, min(n, s0 + step), batch)]
        pipe.step(bs, int(ts[s0]), out, out_ts, n_out)
        torch.cuda.synchronize()
        got.append(pipe.ff.results_to_host(out, out_ts, n_out)[0])
        surv, _, _ = O.map_filter_tuple64(t[s0:s0 + step], ts[s0:s0 + step], 1, 2, 1.0000001, 1)
        exp.append(go.process_batch(O.lift_tuple64(surv), int(ts[s0]))[0])
    pipe.flush(out, out_ts, n_out)
    torch.cuda.synchronize()
    got.append(pipe.ff.results_to_host(out, out_ts, n_out)[0])
    g = O.sort_results(np.concatenate(got)); e = O.sort_results(np.concatenate(exp))
    assert len(g) == len(e) > 0
    assert np.array_equal(g["key"], e["key"]) and np.array_equal(g["id"], e["id"]) and np.array_equal(g["isum"], e["isum"])
    assert np.allclose(g["fsum"], e["fsum"], rtol=1e-6, atol=0)
    assert np.array_equal(g["ts"], e["ts"]) if "ts" in g.dtype.names and "ts" in e.dtype.names else True


def test_mg_pipeline_c_abi_world1(wfb, oracle):
    """The same pipeline with the whole step under the C ABI (wfb_mg_step / wfb_mg_flush): one rank, so the all-to-all is a device copy,
    everything else -- source-side partition by (destination, bucket), size bookkeeping on the communication stream, results three steps late -- is the
    code every rank runs. (bench.py's check covers the NCCL exchange itself at N > 1.)"""
    import torch
    from windflow_b200 import multigpu
    O, ops = oracle, wfb
    win, slide, nb, nkeys, n, batch = 64, 16, 2, 40, 60000, 4096
    t, ts = O.gen_tuple64(0, n, O.KEY_UNIFORM, nkeys)
    f = ops.functors(map_kind=1, iadd=2, fscale=1.0000001, filt_kind=1)
    pipe = multigpu.KeyShardedPipelineC(ops, f, win, slide, nb, 64, 0, 1, torch.device("cuda", 0))
    go = O.FfatGpuOracle(win, slide, nb)
    cap = pipe.max_results(n)
    out = torch.empty(cap * 32, dtype=torch.uint8, device="cuda"); out_ts = torch.empty(cap, dtype=torch.int64, device="cuda")
    n_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    got, exp = [], []
    step = 3 * batch
    for s0 in range(0, n, step):
        bs = [ops.DeviceBatch.from_host(t[b:b + batch], ts[b:b + batch]) for b in range(s0, min(n, s0 + step), batch)]
        pipe.step(bs, int(ts[s0]), out, out_ts, n_out)
        torch.cuda.synchronize()
        got.append(pipe.results_to_host(out, out_ts, n_out)[0])
        surv, _, _ = O.map_filter_tuple64(t[s0:s0 + step], ts[s0:s0 + step], 1, 2, 1.0000001, 1)
        exp.append(go.process_batch(O.lift_tuple64(surv), int(ts[s0]))[0])
    pipe.flush(out, out_ts, n_out)
    torch.cuda.synchronize()
    got.append(pipe.results_to_host(out, out_ts, n_out)[0])
    assert pipe.stats()[1] == 0 and pipe.results_total() == sum(len(x) for x in got)
    g = O.sort_results(np.concatenate(got)); e = O.sort_results(np.concatenate(exp))
    assert len(g) == len(e) > 0
    assert np.array_equal(g["key"], e["key"]) and np.array_equal(g["id"], e["id"]) and np.array_equal(g["isum"], e["isum"])
    assert np.allclose(g["fsum"], e["fsum"], rtol=1e-6, atol=0)
