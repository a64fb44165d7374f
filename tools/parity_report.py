#!/usr/bin/env python
"""Parity report on a B200: CUDA path vs the C oracle (and cv2 when importable) on BASELINE.json's configs at full size,
with the faithful arithmetic and with the library default (relaxed where validated); the relaxed result is also compared
bit for bit with the oracle's restatement of the same relaxation (twref_set_relax(144)).  Writes one JSON line per case.  usage: tools/parity_report.py [--quick]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tidalwave_b200 as tw
from oracle.oracle import FlowParam, RefOracle, cv2_flow, sample_numpy

quick = "--quick" in sys.argv
O = RefOracle()
cases = [
    ("cfg2 S 1920x1080 default", ("S", 1920, 1080, 2, False), dict()),
    ("cfg2 T 1920x1080 default", ("T", 1920, 1080, 1, False), dict()),
    ("cfg2 S+defect 1920x1080 default", ("S", 1920, 1080, 100, True), dict()),
    ("cfg4 S 1280x2000 n5 s1.1 w15 box", ("S", 1280, 2000, 3, False), dict(polyN=5, polySigma=1.1, winSize=15, flags=0)),
    ("cfg4 T 1280x2000 n5 s1.1 w15 box", ("T", 1280, 2000, 6, False), dict(polyN=5, polySigma=1.1, winSize=15, flags=0)),
    ("cfg3 T 3840x2160 lv5 it5", ("T", 3840, 2160, 4, False), dict(pyrLevels=5, pyrIterations=5)),
    ("cfg3 S 3840x2160 lv5 it5", ("S", 3840, 2160, 5, False), dict(pyrLevels=5, pyrIterations=5)),
]
if quick:
    cases = cases[:1] + cases[3:4]
of = tw.OpticalFlow(0, 3840, 2160, 1)
for name, (kind, W, H, seed, defect), kw in cases:
    a, b = tw.synth.make_pair(kind, W, H, seed, defect)
    t0 = time.time(); ref = O.farneback(a, b, FlowParam(**kw)); t_or = time.time() - t0
    try:
        cvf = cv2_flow(a, b, FlowParam(**kw))
    except Exception:
        cvf = None
    row = dict(case=name, oracle_s=round(t_or, 2))
    for mode in ("faithful", "relaxed"):
        of.set_option("arithmetic", int(mode == "relaxed"))
        if mode == "relaxed" and of.arithmetic_in_effect(tw.OpticalFlowParameter(**kw)) != "relaxed":
            continue
        rc, fx, fy, sec = of.calculateInternal(a, b, tw.OpticalFlowParameter(**kw))
        assert rc == 0, of.last_error()
        fl = np.stack([fx, fy], -1)
        d = np.abs(fl - ref)
        st = dict(max=float(d.max()), rms=float(np.sqrt((d ** 2).mean())), frac_gt_1e2=float((d.max(-1) > 1e-2).mean()),
                  bit_equal=float((fl == ref).mean()), gpu_ms=round(sec * 1e3, 3),
                  status_same=sample_numpy(fl)[0] == O.sample(ref)[0],
                  vectors_same=[(v[0], v[1]) for v in sample_numpy(fl)[1]] == [(v[0], v[1]) for v in O.sample(ref)[1]])
        if mode == "relaxed":
            O.set_relax(144)
            rel = O.farneback(a, b, FlowParam(**kw))
            O.set_relax(0)
            st["bit_equal_relaxed_oracle"] = float((fl == rel).mean())
        if cvf is not None:
            dc = np.abs(fl - cvf)
            st["vs_cv2_max"] = float(dc.max()); st["vs_cv2_rms"] = float(np.sqrt((dc ** 2).mean()))
            st["cv2_status_same"] = sample_numpy(fl)[0] == sample_numpy(cvf)[0]
        row[mode] = st
    of.set_option("arithmetic", 1)
    print(json.dumps(row), flush=True)
