#!/usr/bin/env python
"""CPU study (no GPU): how far do the relaxed-arithmetic variants of the oracle (oracle/farneback_ref.c, twref_set_relax)
move the flow away from the faithful oracle and from cv2?  One JSON line per (case, bits).
usage: tools/relax_study.py [--quick] [--4k]"""
import json, os, sys, time
import ctypes as C
import numpy as np
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    ("fixture s2 rev2", ("fixture", "s2"), dict()),
    ("cfg2 S 1920x1080 default", ("S", 1920, 1080, 2, False), dict()),
    ("cfg2 T 1920x1080 default", ("T", 1920, 1080, 1, False), dict()),
    ("cfg2 S+defect 1920x1080 default", ("S", 1920, 1080, 100, True), dict()),
    ("pool S seed 101", ("S", 1920, 1080, 101, False), dict()),
    ("pool S seed 104 defect", ("S", 1920, 1080, 104, True), dict()),
    ("S 1280x2000 n5 s1.1 w15 GAUSS", ("S", 1280, 2000, 3, False), dict(polyN=5, polySigma=1.1, winSize=15)),
]
CASES_4K = [
    ("cfg3 T 3840x2160 lv5 it5", ("T", 3840, 2160, 4, False), dict(pyrLevels=5, pyrIterations=5)),
    ("cfg3 S 3840x2160 lv5 it5", ("S", 3840, 2160, 5, False), dict(pyrLevels=5, pyrIterations=5)),
]
BITS = [int(b) for b in os.environ.get("RELAX_BITS", "1,2,4,3,7").split(",")]


def run(job):
    name, spec, kw, bits = job
    import importlib.util
    spec_ = importlib.util.spec_from_file_location("tw_synth", os.path.join(ROOT, "tidal-wave_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec_); spec_.loader.exec_module(synth)
    from oracle.oracle import FlowParam, RefOracle, cv2_flow, sample_numpy
    if spec[0] == "fixture":
        a = np.load(os.path.join(ROOT, "tests/golden/fixture_s2_expected.npy"))
        b = np.load(os.path.join(ROOT, "tests/golden/fixture_s2_revision2.npy"))
    else:
        a, b = synth.make_pair(*spec)
    O = RefOracle()
    p = FlowParam(**kw)
    O.lib.twref_set_relax(0)
    ref = O.farneback(a, b, p)
    out = []
    cvf = cv2_flow(a, b, p)
    d = np.abs(ref - cvf)
    out.append(dict(case=name, bits=0, vs_cv2_max=float(d.max()), vs_cv2_rms=float(np.sqrt((d ** 2).mean()))))
    for bt in bits:
        O.lib.twref_set_relax(bt)
        fl = O.farneback(a, b, p)
        O.lib.twref_set_relax(0)
        d = np.abs(fl - ref); dc = np.abs(fl - cvf)
        out.append(dict(case=name, bits=bt, max=float(d.max()), rms=float(np.sqrt((d ** 2).mean())),
                        frac_gt_1e3=float((d.max(-1) > 1e-3).mean()), vs_cv2_max=float(dc.max()),
                        vs_cv2_rms=float(np.sqrt((dc ** 2).mean())),
                        status_same=sample_numpy(fl)[0] == sample_numpy(cvf)[0],
                        vectors_same=[(v[0], v[1]) for v in sample_numpy(fl)[1]] == [(v[0], v[1]) for v in sample_numpy(cvf)[1]]))
    return out


if __name__ == "__main__":
    cases = CASES[:3] if "--quick" in sys.argv else CASES
    if "--4k" in sys.argv:
        cases = cases + CASES_4K
    from oracle.oracle import build
    build()
    jobs = [(n, s, k, BITS) for n, s, k in cases]
    with ProcessPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
        for rows in ex.map(run, jobs):
            for r in rows:
                print(json.dumps(r), flush=True)
