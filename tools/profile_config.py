#!/usr/bin/env python
"""Per-kernel-family event timings for any option set / size (B200).  usage: tools/profile_config.py W H batch [key=value ...]"""
import json, sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tidalwave_b200 as tw
W, H, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kw = {}
for a in sys.argv[4:]:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v else int(v)
pairs = [tw.synth.make_pair("S" if i % 2 == 0 else "T", W, H, 200 + i) for i in range(min(B, 4))]
of = tw.OpticalFlow(0, W, H, B)
p = tw.OpticalFlowParameter(**kw)
batch = [pairs[i % len(pairs)] for i in range(B)]
of.calculate_batch(batch, p)          # warm-up + plan
of.profile(True)
n = 5
for _ in range(n):
    of.calculate_batch(batch, p)
prof = of.profile_read()
tot = sum(v["ms"] for v in prof.values())
print(json.dumps({"W": W, "H": H, "batch": B, "opts": kw, "ms_per_pair": tot / n / B,
                  "families_us_per_pair": {k: round(v["ms"] / n / B * 1e3, 1) for k, v in prof.items() if v["launches"]}}))
