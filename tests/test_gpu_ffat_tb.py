"""GPU parity tests (-m gpu): Ffat_Windows_GPU with time-based windows through the C ABI (wfb_ffat_process_tb) against the
oracle's restatement of Ffat_Replica_GPU::process_batch_tb / process_wins_tb and PendingPanes_Queue
(wf/ffat_replica_gpu.hpp:263-420, :870-1047). Bit-exact on keys, window ids, integer aggregates and result timestamps;
floating-point aggregates within 1e-6 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FP_RTOL = 1e-6


def _stream(O, n, nkeys, mode, seed):
    """tuples + timestamps: 'mono' ts = i (the reference's tests), 'jitter' locally out of order, 'gaps' with idle periods
    (empty panes), 'late' with tuples far behind the watermark (dropped once their pane was consumed)."""
    t, _ = O.gen_tuple64(seed, n, O.KEY_RR, nkeys)   # round-robin keys: every key in every batch (win_common_gpu.hpp:104-107)
    rng = np.random.default_rng(seed)
    ts = np.arange(n, dtype=np.int64) * 3
    if mode == "jitter":
        ts = ts + rng.integers(-40, 41, n)
    elif mode == "gaps":
        ts = ts + (np.arange(n) // 700) * 900
    elif mode == "late":
        late = rng.random(n) < 0.03
        ts = np.where(late, ts - rng.integers(200, 3000, n), ts)
    return t, np.maximum(ts, 0).astype(np.uint64)


CASES = [  # win, slide, lateness, nb, nkeys, n, batch, mode, dense
    (40, 10, 0, 1, 4, 6000, 500, "mono", True),
    (40, 10, 0, 3, 5, 8000, 777, "mono", False),
    (64, 16, 0, 2, 7, 9000, 1000, "jitter", True),
    (64, 16, 100, 2, 7, 9000, 1000, "jitter", False),
    (30, 45, 0, 2, 3, 6000, 400, "mono", True),        # hopping (slide > win), pane 15
    (50, 50, 0, 1, 6, 6000, 512, "gaps", True),        # tumbling, idle periods -> empty panes
    (96, 32, 64, 4, 9, 12000, 1500, "late", False),
    (4096, 64, 0, 5, 16, 60000, 8192, "mono", True),   # the bench geometry in microseconds
]


@pytest.mark.parametrize("case", CASES, ids=[f"w{c[0]}_s{c[1]}_l{c[2]}_nb{c[3]}_{c[7]}" for c in CASES])
def test_ffat_tb_vs_oracle(wfb, oracle, case):
    import torch
    O, ops = oracle, wfb
    win, slide, lateness, nb, nkeys, n, batch, mode, dense = case
    t, ts = _stream(O, n, nkeys, mode, 11)
    ff = ops.FfatWindowsGPU(ops.PROG_TUPLE64, win, slide, nb, max_keys=max(nkeys, 8), dense_keys=dense, win_type=1, lateness=lateness)
    tb = O.FfatTbOracle(win, slide, lateness, nb)
    got, gts, exp, ets = [], [], [], []
    for b in range(0, n, batch):
        tb_, tsb = t[b:b + batch], ts[b:b + batch]
        wm = int(tsb.min()) if mode in ("jitter", "late") else int(tsb[0])  # a watermark never exceeds a later timestamp... of on-time tuples
        r, rt = tb.process_batch(O.lift_tuple64(tb_), tsb, wm)
        exp.append(r); ets.append(rt)
        out, out_ts, n_out = ff.process([ops.DeviceBatch.from_host(tb_, tsb, watermark=wm)])
        torch.cuda.synchronize()
        g_, gt_ = ff.results_to_host(out, out_ts, n_out)
        got.append(g_); gts.append(gt_)
    g, gt = O.sort_results(np.concatenate(got), np.concatenate(gts))
    e, et = O.sort_results(np.concatenate(exp), np.concatenate(ets))
    assert len(g) == len(e) > 0, (len(g), len(e))
    assert np.array_equal(g["key"], e["key"]) and np.array_equal(g["id"], e["id"])
    assert np.array_equal(gt, et)
    assert np.array_equal(g["isum"], e["isum"])
    assert np.allclose(g["fsum"], e["fsum"], rtol=FP_RTOL, atol=0)
    assert ff.stats()[1] == 0


def test_ffat_tb_fused_map_filter_and_test_schema(wfb, oracle):
    """Map -> Filter fused in front of the time-based windows, and the reference tests' {key, id, value} schema."""
    import torch
    O, ops = oracle, wfb
    win, slide, nb, nkeys, n, batch = 60, 20, 2, 5, 8000, 640
    t, ts = _stream(O, n, nkeys, "mono", 5)
    f = ops.functors(map_kind=1, iadd=2, fscale=1.0000001, filt_kind=1)
    ff = ops.FfatWindowsGPU(ops.PROG_TUPLE64, win, slide, nb, max_keys=8, dense_keys=True, win_type=1)
    tb = O.FfatTbOracle(win, slide, 0, nb)
    w24 = ops.FfatWindowsGPU(ops.PROG_WFWIN24, win, slide, nb, max_keys=8, win_type=1)
    tb24 = O.FfatTbOracle(win, slide, 0, nb)
    got, exp, got24, exp24 = [], [], [], []
    for b in range(0, n, batch):
        tb_, tsb = t[b:b + batch], ts[b:b + batch]
        wm = int(tsb[0])
        surv, sts, _ = O.map_filter_tuple64(tb_, tsb, 1, 2, 1.0000001, 1)
        exp.append(tb.process_batch(O.lift_tuple64(surv), sts, wm)[0])
        out, out_ts, n_out = ff.process([ops.DeviceBatch.from_host(tb_, tsb, watermark=wm)], pre=f)
        torch.cuda.synchronize()
        got.append(ff.results_to_host(out, out_ts, n_out)[0])
        # {key, id, value}: value = ivalue, windows sum the values
        w = np.zeros(len(tb_), dtype=ops.WFWIN24); w["key"], w["id"], w["value"] = tb_["key"], tb_["id"], tb_["ivalue"]
        lifted = O.lift_tuple64(tb_); lifted["fsum"] = 0.0
        exp24.append(tb24.process_batch(lifted, tsb, wm)[0])
        out, out_ts, n_out = w24.process([ops.DeviceBatch.from_host(w, tsb, watermark=wm)])
        torch.cuda.synchronize()
        got24.append(w24.results_to_host(out, out_ts, n_out)[0])
    g = O.sort_results(np.concatenate(got)); e = O.sort_results(np.concatenate(exp))
    assert len(g) == len(e) > 0 and np.array_equal(g["key"], e["key"]) and np.array_equal(g["id"], e["id"]) and np.array_equal(g["isum"], e["isum"])
    assert np.allclose(g["fsum"], e["fsum"], rtol=FP_RTOL, atol=0)
    g24 = np.concatenate(got24); e24 = O.sort_results(np.concatenate(exp24))
    o = np.lexsort((g24["id"], g24["key"])); g24 = g24[o]
    assert len(g24) == len(e24) > 0 and np.array_equal(g24["key"], e24["key"]) and np.array_equal(g24["id"], e24["id"])
    assert np.array_equal(g24["value"], e24["isum"])


def test_ffat_tb_rejects_wrong_calls(wfb):
    import ctypes as C
    import torch
    ops = wfb
    ff = ops.FfatWindowsGPU(ops.PROG_TUPLE64, 40, 10, 1, max_keys=8, dense_keys=True, win_type=1)
    t = torch.zeros(64 * 10, dtype=torch.uint8, device="cuda")
    with pytest.raises(Exception):
        ff.process([ops.DeviceBatch(t, None, 10, 0)])  # no timestamps
    cb = ops.FfatWindowsGPU(ops.PROG_TUPLE64, 40, 10, 1, max_keys=8, dense_keys=True)
    cb.win_type = 1
    with pytest.raises(Exception):
        cb.process([ops.DeviceBatch(t, torch.zeros(10, dtype=torch.int64, device="cuda"), 10, 0)])  # count-based handle
