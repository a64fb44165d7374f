#!/usr/bin/env python
"""CPU study (no GPU) of relaxed-arithmetic bit sets on the cases tests/test_gpu_relaxed.py holds to the tolerance (the reference
fixtures s1 / s2r2 and the seeded synthetic pairs, default and 5-level / 5-iteration options) plus three 1920x1080 pairs:
distance of oracle(bits) to the faithful oracle, status / vector identity.  usage: tools/relax_cases.py 17,81,144"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from concurrent.futures import ProcessPoolExecutor
def run(job):
    name, opt, bits = job
    import importlib.util
    sp = importlib.util.spec_from_file_location("tw_synth", os.path.join(ROOT, "tidal-wave_b200", "synth.py"))
    synth = importlib.util.module_from_spec(sp); sp.loader.exec_module(synth)
    from oracle.oracle import FlowParam, RefOracle, sample_numpy
    OPTS = {"default": dict(), "cfg3": dict(pyrLevels=5, pyrIterations=5)}
    G = os.path.join(ROOT, "tests", "golden") + "/"
    if name == "s1": a = np.load(G + "fixture_s1_expected.npy"); b = a
    elif name == "s2r2": a = np.load(G + "fixture_s2_expected.npy"); b = np.load(G + "fixture_s2_revision2.npy")
    elif name == "S256": a, b = synth.make_pair("S", 256, 192, 2, False)
    elif name == "T256": a, b = synth.make_pair("T", 256, 160, 1, False)
    elif name == "Sdef": a, b = synth.make_pair("S", 320, 200, 8, True)
    elif name == "S1080": a, b = synth.make_pair("S", 1920, 1080, 100, True)
    elif name == "S1080b": a, b = synth.make_pair("S", 1920, 1080, 104, True)
    elif name == "T1080": a, b = synth.make_pair("T", 1920, 1080, 1, False)
    O = RefOracle(); p = FlowParam(**OPTS[opt])
    O.set_relax(0); ref = O.farneback(a, b, p)
    out = []
    for bt in bits:
        O.set_relax(bt); fl = O.farneback(a, b, p); O.set_relax(0)
        d = np.abs(fl - ref)
        out.append(dict(case=name, opt=opt, bits=bt, max=float(d.max()), rms=float(np.sqrt((d**2).mean())),
                        status_same=sample_numpy(fl)[0] == sample_numpy(ref)[0],
                        vec_same=[(v[0], v[1]) for v in sample_numpy(fl)[1]] == [(v[0], v[1]) for v in sample_numpy(ref)[1]]))
    return out
if __name__ == "__main__":
    bits = [int(x) for x in sys.argv[1].split(",")]
    jobs = [(n, o, bits) for n in ("s1", "s2r2", "S256", "T256", "Sdef") for o in ("default", "cfg3")] + [(n, "default", bits) for n in ("S1080", "S1080b", "T1080")]
    with ProcessPoolExecutor(8) as ex:
        for rows in ex.map(run, jobs):
            for r in rows: print(json.dumps(r), flush=True)
