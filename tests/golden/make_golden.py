"""Generates tests/golden/*.npz from the REFERENCE itself (run in the build container, where /root/reference exists).

The reference ships no golden vectors for this path, so these are produced by running the reference's
own wf/flatfat.hpp (compiled unmodified into oracle/_ref/libwfref_flatfat.so by oracle/Makefile) under the restated
FFAT_Replica count-based trigger loop (oracle/ref_flatfat.cpp). Each file holds the lifted input stream and the
windows the reference emits (complete windows, then the end-of-stream flush of the CPU operator).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CONFIGS = [  # (name, win, slide, nkeys, n, batch)
    ("cb_w4_s2_k3", 4, 2, 3, 3000, 257),
    ("cb_w10_s3_k5", 10, 3, 5, 4000, 100),
    ("cb_w16_s16_k2", 16, 16, 2, 2000, 64),
    ("cb_w64_s16_k7", 64, 16, 7, 6000, 333),
    ("cb_w1024_s32_k4", 1024, 32, 4, 12000, 1000),
]


def main():
    assert O.ref_cpu_lib() is not None, "oracle/_ref/libwfref_flatfat.so missing: run make -C oracle"
    for name, win, slide, nkeys, n, batch in CONFIGS:
        rng = np.random.default_rng(sum(name.encode()) * 7919)
        res = np.zeros(n, dtype=O.RES)
        res["key"] = rng.integers(0, nkeys, n)
        res["isum"] = rng.integers(-1000, 1000, n)
        res["fsum"] = rng.random(n)
        ref = O.RefFfatCpu(win, slide)
        outs, tss = [], []
        for b in range(0, n, batch):
            o, t = ref.process(res[b:b + batch], b)
            outs.append(o); tss.append(t)
        eo, et = ref.eos()
        eo, et = O.sort_results(eo, et)
        out = np.concatenate(outs); ts = np.concatenate(tss)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), win=win, slide=slide, batch=batch,
                            key=res["key"].astype(np.uint16), isum=res["isum"].astype(np.int16), fsum=res["fsum"],
                            out_key=out["key"].astype(np.uint16), out_id=out["id"].astype(np.uint32), out_isum=out["isum"],
                            out_fsum=out["fsum"], out_ts=ts,
                            eos_key=eo["key"].astype(np.uint16), eos_id=eo["id"].astype(np.uint32), eos_isum=eo["isum"],
                            eos_fsum=eo["fsum"])
        print(name, "windows", len(out), "eos", len(eo))


if __name__ == "__main__":
    main()
