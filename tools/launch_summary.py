#!/usr/bin/env python
"""Summarises an ncu --csv launch list (gpu__time_duration.sum [+ dram bytes]) per kernel name. usage: launch_summary.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value"); mi = H.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg.setdefault(r[ki][:72], collections.defaultdict(list))[r[mi]].append(float(r[vi].replace(",", "")))
for k, v in agg.items():
    t = v["gpu__time_duration.sum"]
    extra = ""
    if "dram__bytes_read.sum" in v:
        # ncu prints bytes scaled (Mbyte etc.) in the CSV unit column; values here are as printed
        extra = "  dram rd %.1f wr %.1f" % (sum(v["dram__bytes_read.sum"]) / len(t), sum(v["dram__bytes_write.sum"]) / len(t))
    print(f"{k:72s} n={len(t):3d} avg_us={sum(t) / len(t) / 1000:8.1f}{extra}")
