#!/usr/bin/env python
"""Per-CUDA-source-line instruction / stall-sample shares from an .ncu-rep (needs --import-source on and -lineinfo).
usage: ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hi]; ii = H.index("Instructions Executed"); wi = H.index("Warp Stall Sampling (All Samples)")
def num(x):
    try: return float(x)
    except ValueError: return 0.0
lines = [(r[0], r[1], num(r[ii]), num(r[wi])) for r in rows[hi + 1:] if len(r) > wi and r[0].isdigit()]
ti = sum(l[2] for l in lines); ts = sum(l[3] for l in lines)
print(f"total warp instructions {ti:.0f}, stall samples {ts:.0f}")
for ln, src, n, w in lines:
    if n > ti * thr / 100 or w > ts * thr / 100:
        print(f"{ln:>5} instr {n / ti * 100:5.1f}%  stall {w / ts * 100:5.1f}%  {src.strip()[:120]}")
