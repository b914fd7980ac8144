"""CPU test (-m "not gpu") of the multi-GPU host logic with the gloo backend, world size 2: segment ownership, the
count + segment all-to-all, source-rank-order concatenation. The per-rank compute is done by the ORACLE here (this is
a test of the exchange plumbing); the merged windows must equal the single-operator oracle on the whole stream."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, pickle
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["WFB_ROOT"])
    from oracle import oracle as O
    from windflow_b200 import multigpu as M

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    win, slide, nb, nkeys, seg_tuples, steps = 16, 4, 2, 37, 3000, 6
    go = O.FfatGpuOracle(win, slide, nb)
    mine = []
    for t in range(steps):
        first, last = M.owner_span(t, rank, world, seg_tuples)
        tup, ts = O.gen_tuple64(first, last - first, O.KEY_UNIFORM, nkeys)
        surv, sts, _ = O.map_filter_tuple64(tup, ts, 1, 2, 1.0000001, 1)
        dest = O.route(surv["key"], world)
        order = np.argsort(dest, kind="stable")
        part = np.ascontiguousarray(surv[order])
        counts = np.bincount(dest, minlength=world)
        send = torch.from_numpy(part.view(np.uint8).reshape(-1).copy())
        rc, rw = M.exchange_counts(torch.tensor(counts.tolist(), dtype=torch.int64), int(ts[0]))
        assert rw.tolist() == [M.owner_span(t, s, world, seg_tuples)[0] for s in range(world)]
        recv, offs = M.exchange_segments(send, counts.tolist(), rc.tolist(), 64)
        for s in range(world):
            chunk = recv[offs[s] * 64:offs[s + 1] * 64].numpy().view(O.TUPLE64)
            assert (chunk["key"] % world == rank).all()
            assert (np.diff(chunk["id"].astype(np.int64)) > 0).all()        # arrival order kept inside a chunk
            r, _ = go.process_batch(O.lift_tuple64(chunk), int(rw[s]))
            mine.append(r)
    mine = np.concatenate(mine) if mine else np.zeros(0, dtype=O.RES)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.tobytes())
    if rank == 0:
        got = np.concatenate([np.frombuffer(b, dtype=O.RES) for b in gathered])
        # single operator over the whole stream, same segments in global order
        ref = O.FfatGpuOracle(win, slide, nb)
        exp = []
        for t in range(steps):
            for s in range(world):
                first, last = M.owner_span(t, s, world, seg_tuples)
                tup, ts = O.gen_tuple64(first, last - first, O.KEY_UNIFORM, nkeys)
                surv, _, _ = O.map_filter_tuple64(tup, ts, 1, 2, 1.0000001, 1)
                r, _ = ref.process_batch(O.lift_tuple64(surv), int(ts[0]))
                exp.append(r)
        exp = O.sort_results(np.concatenate(exp)); got = O.sort_results(got)
        assert len(got) == len(exp) > 0, (len(got), len(exp))
        assert np.array_equal(got["key"], exp["key"]) and np.array_equal(got["id"], exp["id"])
        assert np.array_equal(got["isum"], exp["isum"]) and np.allclose(got["fsum"], exp["fsum"], rtol=1e-9)
        print("MULTI_OK", len(got))
    dist.destroy_process_group()
''')


def test_keyby_sharded_exchange_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, WFB_ROOT=ROOT, OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() % 400)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "MULTI_OK" in out.stdout


def test_tile_layout():
    """Receive-buffer layout of the in-place window update: chunk s starts at 256 * (tiles of the chunks before it)."""
    from windflow_b200 import multigpu as M
    offs, total = M.tile_layout([0, 1, 256, 257, 0, 1000])
    assert offs == [0, 0, 256, 512, 1024, 1024] and total == 1024 + 1024
    assert M.tile_layout([]) == ([], 0)


def test_bucketed_exchange_model_keeps_per_key_stream_order():
    """numpy model of the bucketed exchange of wfb_mg_step: every source partitions its survivors, stably, on the
    destination-major virtual slot (key % n) * L + key // n shifted to at most 1024 bins; a destination concatenates, bucket after bucket,
    the runs of the sources in rank order and splits every coarse bucket into sub-buckets, stably. Claim the CUDA path relies on: for
    every key the resulting item order is the global stream order (source rank major, arrival order inside a source), and the bucket of
    an item is a function of its slot alone. (The kernels themselves are checked on the GPU: tests/test_gpu_keyed.py, bench.py --check.)"""
    import numpy as np
    rng = np.random.default_rng(7)
    for n, nkeys, per_src in ((2, 65536, 40000), (4, 5000, 30000), (8, 65536, 20000), (3, 1000, 9000)):
        keys_per = (nkeys + n - 1) // n
        L = 1
        while L < keys_per:
            L <<= 1
        span = 1
        while span < L * n:
            span <<= 1
        shift = 0
        while (span >> shift) > 1024:
            shift += 1
        bps = L >> shift
        assert bps >= 1 and n * L <= 65536
        nsub, shift2 = 1, shift
        while nsub * 2 * bps <= 1024 and nsub * 2 <= 8 and shift2 > 0:
            nsub, shift2 = nsub * 2, shift2 - 1
        srcs = [rng.integers(0, nkeys, per_src) for _ in range(n)]                     # surviving keys of every source, arrival order
        glob = [(int(k), s, i) for s in range(n) for i, k in enumerate(srcs[s])]       # global stream order: source rank major
        delivered = {d: [] for d in range(n)}                                          # per destination: runs [(source, bucket, items)]
        for s in range(n):
            k = srcs[s]
            v = (k % n) * L + k // n
            order = np.argsort(v >> shift, kind="stable")                              # the source's ONE stable partition pass
            bins = (v >> shift)[order]
            for d in range(n):
                sel = (bins >= d * bps) & (bins < (d + 1) * bps)
                delivered[d].append((s, bins[sel] - d * bps, v[order][sel] & (L - 1), order[sel]))
        for d in range(n):
            # destination: bucket-major, sub-bucket, source rank, arrival order
            rows = []
            for s, b, slot, idx in delivered[d]:
                assert (np.diff(b) >= 0).all()                                         # a source's runs arrive bucket after bucket
                sub = (slot >> shift2) & (nsub - 1)
                assert ((slot >> shift) == b).all()
                rows.append(np.stack([b * nsub + sub, np.full_like(b, s), np.arange(len(b)), slot, idx], axis=1))
            allr = np.concatenate(rows)
            allr = allr[np.lexsort((allr[:, 2], allr[:, 1], allr[:, 0]))]              # what k_mg_count / scan / split produce
            assert ((allr[:, 3] >> shift2) == allr[:, 0]).all()                        # bucket of the update kernel = slot >> shift2
            assert (1 << shift2) <= 64                                                 # at most 64 slots per bucket (BK_KEYS)
            for slot in np.unique(allr[:, 3])[:200]:
                mine = allr[allr[:, 3] == slot]
                key = int(slot) * n + d
                want = [(s, i) for (k, s, i) in glob if k == key]
                assert [(int(r[1]), int(r[4])) for r in mine] == want
