#!/bin/bash
# run HERE after tools/profile_round.sh came back: details CSV + per-source-line shares of every capture into profiles/, dram bytes into profiles/traffic.json
TAG=${1:-r2}
for k in k_tile_pass k_wide_scatter_ranked k_ffat_update_buckets; do
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page details --csv > profiles/${TAG}_ncu_full_$k.csv 2>/dev/null
  python tools/ncu_lines.py gpurun_out/${TAG}_$k.ncu-rep 1.5 > profiles/${TAG}_ncu_lines_$k.txt 2>/dev/null
done
for f in launches.txt launches_raw.csv mgpath_launches.txt facade_launches.txt; do cp gpurun_out/${TAG}_$f profiles/ 2>/dev/null; done
python - "$TAG" <<'PY'
import csv, json, subprocess, sys
tag = sys.argv[1]
out = {}
names = {"k_tile_pass": "tile_pass (map, filter, lift, key->slot)", "k_wide_scatter_ranked": "partition scatter", "k_ffat_update_buckets": "window update"}
for k, label in names.items():
    rows = list(csv.reader(subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_{k}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    if len(rows) < 3:
        continue
    H, U, V = rows[0], rows[1], rows[2]
    def val(name):
        i = H.index(name); v = float(V[i].replace(",", "")); u = U[i].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
    out[k] = {"dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"), "ncu_time_us": float(V[H.index("gpu__time_duration.sum")].replace(",", "")),
              "source": f"profiles/{tag}_ncu_full_{k}.csv"}
t = out
flat = {"source": f"one ncu --set full capture per kernel of the default bench step (8 388 608 tuples), tools/profile_round.sh {tag}", "kernels": t}
if "k_tile_pass" in t:
    flat["tile_pass (map, filter, lift, key->slot)"] = t["k_tile_pass"]["dram_read_bytes"] + t["k_tile_pass"]["dram_write_bytes"]
if "k_ffat_update_buckets" in t:
    flat["window update + queries"] = t["k_ffat_update_buckets"]["dram_read_bytes"] + t["k_ffat_update_buckets"]["dram_write_bytes"]
if "k_wide_scatter_ranked" in t:
    flat["partition (per-tile counts -> offsets -> scatter)"] = t["k_wide_scatter_ranked"]["dram_read_bytes"] + t["k_wide_scatter_ranked"]["dram_write_bytes"]
if len(t) == 3:
    flat["pipeline_dram_bytes_per_call"] = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in t.values())
json.dump(flat, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(flat, indent=1)[:1500])
PY
