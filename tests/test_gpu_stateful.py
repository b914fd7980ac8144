"""GPU parity tests (-m gpu): keyed-stateful Map_GPU / Filter_GPU through the C ABI (wfb_map_stateful /
wfb_filter_stateful) against the oracle's per-key sequential restatement. Bit-exact (integer state, byte-identical tuples,
stable order of the survivors)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _wftest16(O, ops, start, n, nkeys, mode):
    t, ts = O.gen_tuple64(start, n, mode, nkeys)
    w = np.zeros(n, dtype=ops.WFTEST16)
    w["key"], w["value"] = t["key"], t["ivalue"]
    return w, ts


@pytest.mark.parametrize("nkeys,dense", [(1, True), (7, False), (100, True), (5000, False), (60000, True)])
@pytest.mark.parametrize("kind", [1, 2])
def test_map_stateful_wftest16(wfb, oracle, nkeys, dense, kind):
    import torch
    O, ops = oracle, wfb
    ks = ops.KeyedState(ops.PROG_WFTEST16, max_keys=max(nkeys, 8), dense_keys=dense)
    f = ops.functors(map_kind=kind)
    state = {}
    start = 0
    for sizes in ([3000, 1, 0, 257], [4096], [700, 700, 700]):
        hosts, devs = [], []
        for n in sizes:
            w, ts = _wftest16(O, ops, start, n, nkeys, O.KEY_UNIFORM if nkeys > 1 else O.KEY_RR)
            start += n
            hosts.append(w)
            devs.append(ops.DeviceBatch.from_host(w, ts) if n else ops.DeviceBatch(torch.empty(0, dtype=torch.uint8, device="cuda"), None, 0, 0))
        ks.map(devs, f)
        torch.cuda.synchronize()
        for w, d in zip(hosts, devs):
            exp = O.stateful_map(w, "value", state, kind)
            if len(w):
                got = ops.to_host(d.tuples, ops.WFTEST16)
                assert got.tobytes() == exp.tobytes()


@pytest.mark.parametrize("filt", [(0, 1), (1, 1), (2, 3)])
def test_filter_stateful_tuple64(wfb, oracle, filt):
    import torch
    O, ops = oracle, wfb
    kind, mod = filt
    nkeys = 300
    ks = ops.KeyedState(ops.PROG_TUPLE64, max_keys=512, dense_keys=False)
    f = ops.functors(filt_kind=kind, mod=mod)
    state, start = {}, 0
    for sizes in ([5000, 3, 0, 1025], [65536], [100] * 5):
        hosts, ins, outs = [], [], []
        for n in sizes:
            t, ts = O.gen_tuple64(start, n, O.KEY_UNIFORM, nkeys)
            t["pad"] = np.arange(n * 4, dtype=np.uint64).reshape(n, 4) + start
            start += n
            hosts.append((t, ts))
            b = ops.DeviceBatch.from_host(t, ts) if n else ops.DeviceBatch(torch.empty(0, dtype=torch.uint8, device="cuda"), torch.empty(0, dtype=torch.int64, device="cuda"), 0, 0)
            ins.append(b)
            outs.append(ops.DeviceBatch(torch.empty_like(b.tuples), torch.empty_like(b.ts), n, 0))
        n_out = torch.full((len(sizes),), 99, dtype=torch.int32, device="cuda")
        ks.filter(ins, f, outs, n_out)
        torch.cuda.synchronize()
        no = n_out.cpu().numpy()
        for i, (t, ts) in enumerate(hosts):
            exp, ets, _ = O.stateful_filter(t, ts, "ivalue", state, kind, mod)
            assert no[i] == len(exp), (i, no[i], len(exp))
            if len(exp):
                got = ops.to_host(outs[i].tuples, ops.TUPLE64)[:len(exp)]
                assert got.tobytes() == exp.tobytes()
                assert np.array_equal(ops.ts_to_host(outs[i].ts)[:len(exp)], ets)


def test_stateful_unsupported_program(wfb):
    ops = wfb
    with pytest.raises(Exception):
        ops.KeyedState(ops.PROG_LIFTED32, max_keys=8)
