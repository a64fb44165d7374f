#!/usr/bin/env python
"""Device-resident timing of one BASELINE config with the per-family breakdown (the `configs` leg of bench.py on its own).
usage: tools/cfg_bench.py cfg3|cfg4|cfg2 [batch] [option=value ...]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import tidalwave_b200 as tw
key = sys.argv[1]
cfg = {"cfg2": (1920, 1080, 16, dict(), [("S", 3), ("T", 6)] * 8),
       "cfg3": (3840, 2160, 8, dict(pyrLevels=5, pyrIterations=5), [("T", 4), ("S", 5), ("T", 14), ("S", 15)]),
       "cfg4": (1280, 2000, 32, dict(polyN=5, polySigma=1.1, winSize=15, flags=0), [("S", 3), ("T", 6)])}[key]
w, h, b, kw, gen = cfg
args = [a for a in sys.argv[2:] if "=" not in a]
if args: b = int(args[0])
made = {ks: tw.synth.make_pair(ks[0], w, h, ks[1], False) for ks in dict.fromkeys(gen)}
prs = [made[ks] for ks in (gen * b)[:b]]
lib = tw.load(); dist = tw.dist.Dist(); p = tw.OpticalFlowParameter(**kw)
R = bench.Resident(tw, lib, 0, prs, p, w, h)
for a in sys.argv[2:]:
    if "=" in a:
        k, v = a.split("="); R.of.set_option(k, int(v))
ms = R.timed(20, 3, dist); v = b * 20 / (ms * 1e-3)
R.of.profile(True); R.timed(5, 0, dist)
pf = {k: v_ for k, v_ in R.of.profile_read().items() if v_["launches"] > 0}
peak, _ = bench.measured_peaks()
print(json.dumps({"cfg": key, "batch": b, "pairs_per_s": round(v, 1), "ms_per_pair": round(ms / 20 / b, 4), "frac": round(v * bench.B_ALG[key] / 1e9 / peak, 3),
                  "families_ms_per_pair": {k: round(v_["ms"] / 5 / b, 4) for k, v_ in pf.items()}}))
R.close()
