"""GPU parity tests (-m gpu): Map_GPU / Filter_GPU through the C ABI against the oracle. Bit-exact (integer masks,
stable order, byte-identical survivors)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 31, 255, 256, 257, 1000, 4097, 65536, 200003]


@pytest.mark.parametrize("n", SIZES)
def test_map_tuple64(wfb, oracle, n):
    import torch
    O, ops = oracle, wfb
    t, ts = O.gen_tuple64(123, n, O.KEY_UNIFORM, 1000)
    t["pad"] = np.arange(n * 4, dtype=np.uint64).reshape(n, 4)  # payload must survive byte for byte
    eng = ops.Engine(ops.PROG_TUPLE64)
    b = ops.DeviceBatch.from_host(t, ts)
    eng.map(b, ops.functors(map_kind=1, iadd=2, fscale=1.0000001))
    torch.cuda.synchronize()
    got = ops.to_host(b.tuples, ops.TUPLE64)
    exp = t.copy()
    exp["ivalue"] += 2
    exp["fvalue"] *= 1.0000001
    assert got.tobytes() == exp.tobytes()


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("filt", [(1, 1), (2, 3), (0, 1)])
def test_map_filter_tuple64(wfb, oracle, n, filt):
    import torch
    O, ops = oracle, wfb
    kind, mod = filt
    t, ts = O.gen_tuple64(77, n, O.KEY_UNIFORM, 1000)
    t["pad"] = np.arange(n * 4, dtype=np.uint64).reshape(n, 4)
    eng = ops.Engine(ops.PROG_TUPLE64)
    b = ops.DeviceBatch.from_host(t, ts)
    out, n_out = eng.map_filter(b, ops.functors(map_kind=1, iadd=2, fscale=1.0000001, filt_kind=kind, mod=mod))
    torch.cuda.synchronize()
    exp, exp_ts, mask = O.map_filter_tuple64(t, ts, 1, 2, 1.0000001, kind, mod)
    k = int(n_out.item())
    assert k == len(exp) == int(mask.sum())
    assert ops.to_host(out.tuples, ops.TUPLE64)[:k].tobytes() == exp.tobytes()
    assert np.array_equal(ops.ts_to_host(out.ts)[:k], exp_ts)
    # input untouched when out != in
    assert ops.to_host(b.tuples, ops.TUPLE64).tobytes() == t.tobytes()


def test_filter_in_place_and_all_dropped(wfb, oracle):
    import torch
    O, ops = oracle, wfb
    n = 70001
    t, ts = O.gen_tuple64(5, n, O.KEY_UNIFORM, 1000)
    eng = ops.Engine(ops.PROG_TUPLE64)
    b = ops.DeviceBatch.from_host(t, ts)
    out, n_out = eng.map_filter(b, ops.functors(filt_kind=1), out=b)
    torch.cuda.synchronize()
    exp, exp_ts, _ = O.map_filter_tuple64(t, ts, 0, 0, 1.0, 1)
    k = int(n_out.item())
    assert k == len(exp)
    assert ops.to_host(b.tuples, ops.TUPLE64)[:k].tobytes() == exp.tobytes()
    assert np.array_equal(ops.ts_to_host(b.ts)[:k], exp_ts)
    # everything dropped: ivalue in [0, 65535] is never a multiple of 2^20
    b2 = ops.DeviceBatch.from_host(t, ts)
    t0 = t.copy(); t0["ivalue"] |= 1
    b2 = ops.DeviceBatch.from_host(t0, ts)
    out, n_out = eng.map_filter(b2, ops.functors(filt_kind=1))
    torch.cuda.synchronize()
    assert int(n_out.item()) == 0
    # empty batch
    e = ops.DeviceBatch(torch.empty(0, dtype=torch.uint8, device="cuda"), torch.empty(0, dtype=torch.int64, device="cuda"), 0)
    out, n_out = eng.map_filter(e, ops.functors(filt_kind=1))
    torch.cuda.synchronize()
    assert int(n_out.item()) == 0


def test_many_launches_one_engine(wfb, oracle):
    """Epoch-tagged tile states and the running ticket counter across many launches / sizes."""
    import torch
    O, ops = oracle, wfb
    eng = ops.Engine(ops.PROG_TUPLE64)
    rng = np.random.default_rng(0)
    for it in range(40):
        n = int(rng.integers(1, 30000))
        t, ts = O.gen_tuple64(it * 1000, n, O.KEY_UNIFORM, 100)
        b = ops.DeviceBatch.from_host(t, ts)
        out, n_out = eng.map_filter(b, ops.functors(map_kind=1, iadd=it, fscale=1.0, filt_kind=2, mod=3))
        exp, exp_ts, _ = O.map_filter_tuple64(t, ts, 1, it, 1.0, 2, 3)
        k = int(n_out.item())
        assert k == len(exp)
        assert ops.to_host(out.tuples, ops.TUPLE64)[:k].tobytes() == exp.tobytes()
    assert eng.launches == 40


@pytest.mark.parametrize("prog", ["wftest16", "wfwin24"])
@pytest.mark.parametrize("n", [1, 3, 255, 257, 1001, 50000])
def test_reference_test_schemas(wfb, oracle, prog, n):
    """The reference tests' own functors: value + 2 (graph_common_gpu.hpp:245-253), value % mod == 0 (:198-215)."""
    import torch
    O, ops = oracle, wfb
    rng = np.random.default_rng(n)
    if prog == "wftest16":
        pid, dt = ops.PROG_WFTEST16, ops.WFTEST16
    else:
        pid, dt = ops.PROG_WFWIN24, ops.WFWIN24
    t = np.zeros(n, dtype=dt)
    t["key"] = rng.integers(0, 9, n)
    t["value"] = rng.integers(-50, 1000, n)
    if prog == "wfwin24":
        t["id"] = np.arange(n)
    ts = np.arange(n, dtype=np.uint64) * 3
    eng = ops.Engine(pid)
    b = ops.DeviceBatch.from_host(t, ts)
    out, n_out = eng.map_filter(b, ops.functors(map_kind=1, iadd=2, filt_kind=2, mod=4))
    torch.cuda.synchronize()
    v = t["value"] + 2
    m = O.filter_mask(v, O.FILT_MOD, 4)
    exp = t.copy(); exp["value"] = v; exp = exp[m]
    k = int(n_out.item())
    assert k == len(exp)
    assert ops.to_host(out.tuples, dt)[:k].tobytes() == exp.tobytes()
    assert np.array_equal(ops.ts_to_host(out.ts)[:k], ts[m])
    eng.map(b, ops.functors(map_kind=1, iadd=2))
    torch.cuda.synchronize()
    e2 = t.copy(); e2["value"] += 2
    assert ops.to_host(b.tuples, dt).tobytes() == e2.tobytes()


def test_device_generator_matches_host(wfb, oracle):
    import torch
    O, ops = oracle, wfb
    for mode, nk in [(O.KEY_RR, 100), (O.KEY_UNIFORM, 65536)]:
        b = ops.gen_tuple64(999, 10000, mode, nk)
        torch.cuda.synchronize()
        t, ts = O.gen_tuple64(999, 10000, mode, nk)
        assert ops.to_host(b.tuples, ops.TUPLE64).tobytes() == t.tobytes()
        assert np.array_equal(ops.ts_to_host(b.ts), ts)
    cdf = O.zipf_cdf(5000)
    dcdf = torch.from_numpy(cdf).cuda()
    b = ops.gen_tuple64(0, 20000, O.KEY_ZIPF, 5000, zipf_cdf=dcdf)
    torch.cuda.synchronize()
    t, ts = O.gen_tuple64(0, 20000, O.KEY_ZIPF, 5000, cdf=cdf)
    assert ops.to_host(b.tuples, ops.TUPLE64).tobytes() == t.tobytes()


@pytest.mark.parametrize("inplace", [False, True])
def test_map_filter_batches(wfb, oracle, inplace):
    """K queued batches (ragged sizes, one empty) in one launch == K single-batch calls == the oracle, byte for byte."""
    import torch
    O, ops = oracle, wfb
    sizes = [65536, 1, 0, 257, 4097, 30000, 255, 65536]
    f = ops.functors(map_kind=1, iadd=2, fscale=1.0000001, filt_kind=2, mod=3)
    eng = ops.Engine(ops.PROG_TUPLE64)
    ins, outs, hosts, start = [], [], [], 0
    for n in sizes:
        t, ts = O.gen_tuple64(start, n, O.KEY_UNIFORM, 1000)
        t["pad"] = np.arange(n * 4, dtype=np.uint64).reshape(n, 4) + start
        hosts.append((t, ts)); start += n
        b = ops.DeviceBatch.from_host(t, ts) if n else ops.DeviceBatch(torch.empty(0, dtype=torch.uint8, device="cuda"), torch.empty(0, dtype=torch.int64, device="cuda"), 0, 0)
        ins.append(b)
        outs.append(b if inplace else ops.DeviceBatch(torch.empty_like(b.tuples), torch.empty_like(b.ts), n, 0))
    n_out = torch.full((len(sizes),), 12345, dtype=torch.int32, device="cuda")
    eng.map_filter_batches(ins, f, outs, n_out)
    torch.cuda.synchronize()
    no = n_out.cpu().numpy()
    for i, (t, ts) in enumerate(hosts):
        exp, exp_ts, _ = O.map_filter_tuple64(t, ts, 1, 2, 1.0000001, 2, 3)
        assert no[i] == len(exp)
        got = ops.to_host(outs[i].tuples, ops.TUPLE64, len(exp)) if len(exp) else exp
        assert got.tobytes() == exp.tobytes()
        if len(exp):
            assert np.array_equal(ops.ts_to_host(outs[i].ts, len(exp)), exp_ts)
