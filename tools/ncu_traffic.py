#!/usr/bin/env python
"""DRAM bytes per launch per kernel family from an ncu summary CSV (tools/ncu_summary.py) of ONE bench step
(launch order: fused level images, then per scale: poly-exp, first update, window kernel x pyrIterations, the last of which is
the 'last' iteration).  bench.py reads the result (profiles/*_traffic.json) for roofline.traffic.
usage: tools/ncu_traffic.py summary.csv batch arithmetic(relaxed|faithful) [iterations=3] > profiles/rXX_traffic.json"""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
batch, arithmetic = int(sys.argv[2]), sys.argv[3]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
hdr = rows[0]
col = lambda name: next(i for i, h in enumerate(hdr) if h.startswith(name))
iname, idur, ird, iwr = col("Kernel Name"), col("gpu__time_duration.sum"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
unit = lambda h: h[h.index("[") + 1:h.index("]")]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dscale = {"ns": 1e-3, "us": 1, "ms": 1e3}[unit(hdr[idur])]
fam = {}
g = 0
for r in rows[1:]:
    n = r[iname]
    if "gauss_last_sparse" in n:
        f = "gauss_last"
    elif "gauss_iter" in n or "gauss_strip" in n:
        f = "gauss_last" if g % iters == iters - 1 else "gauss_iter"
        g += 1
    elif "polyexp" in n: f = "polyexp"
    elif "first_update" in n: f = "first_update"
    elif "level_" in n: f = "level_image"
    else: continue
    d = fam.setdefault(f, {"launches": 0, "us_total": 0.0, "bytes": 0.0})
    d["launches"] += 1
    d["us_total"] += float(r[idur]) * dscale
    d["bytes"] += float(r[ird]) * scale[unit(hdr[ird])] + float(r[iwr]) * scale[unit(hdr[iwr])]
out = {"note": "ncu --set full --clock-control none, TW_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e, one step; dram bytes = "
               "dram__bytes_read.sum + dram__bytes_write.sum averaged over the launches of the family; source: " + sys.argv[1],
       "batch": batch, "arithmetic": arithmetic}
for f, d in fam.items():
    out[f] = {"launches": d["launches"], "us_total": round(d["us_total"], 1), "dram_bytes_per_launch": int(d["bytes"] / d["launches"])}
print(json.dumps(out, indent=1))
