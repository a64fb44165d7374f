#!/usr/bin/env python
"""Generates tests/golden/ (run in the build container, where /root/reference and cv2 exist).

* fixture_*.npy      -- the reference's test images decoded exactly as the reference does
                        (cv::imread(path, IMREAD_GRAYSCALE), /root/reference/src/opticalflow.cpp:37,44),
                        via cv2.imread(..., cv2.IMREAD_GRAYSCALE).  revision1/* and revision2/scenario1
                        are pixel-identical to expected/*, so only 3 distinct images are stored.
* reference_golden.json -- the 24 golden vectors + statuses + sizes of /root/reference/test/index.coffee:12-96.
* cv2_flows.npz      -- full-field flows from cv2 4.13.0 calcOpticalFlowFarneback (IPP off, 1 thread) on the
                        fixture pairs and on small seeded synthetic pairs, for the option sets in OPTS.
The GPU box has no /root/reference: tests read only these files.
"""
import importlib.util
import json
import os
import re
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import FlowParam, cv2_flow  # noqa: E402

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "tidal-wave_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

REF = "/root/reference/test"
OUT = os.path.join(ROOT, "tests", "golden")

OPTS = {
    "default": dict(),
    "cfg4": dict(polyN=5, polySigma=1.1, winSize=15, flags=0),
    "cfg3": dict(pyrLevels=5, pyrIterations=5),
    "box31": dict(flags=0),
    "g15n5": dict(polyN=5, polySigma=1.1, winSize=15),
}

SYNTH = {  # name -> (kind, W, H, seed, defect)
    "S256": ("S", 256, 192, 2, False),
    "T256": ("T", 256, 160, 1, False),
    "Sdef": ("S", 320, 200, 8, True),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    fx = {}
    for rev in ("expected", "revision1", "revision2"):
        for sc, fn in (("scenario1", "capture1.jpg"), ("scenario2", "capture2.png")):
            fx[(rev, sc)] = cv2.imread(os.path.join(REF, "fixture", rev, sc, fn), cv2.IMREAD_GRAYSCALE)
    assert np.array_equal(fx[("expected", "scenario1")], fx[("revision1", "scenario1")])
    assert np.array_equal(fx[("expected", "scenario1")], fx[("revision2", "scenario1")])
    assert np.array_equal(fx[("expected", "scenario2")], fx[("revision1", "scenario2")])
    np.save(os.path.join(OUT, "fixture_s1_expected.npy"), fx[("expected", "scenario1")])
    np.save(os.path.join(OUT, "fixture_s2_expected.npy"), fx[("expected", "scenario2")])
    np.save(os.path.join(OUT, "fixture_s2_revision2.npy"), fx[("revision2", "scenario2")])

    # golden vectors, parsed from the reference's own test file (data, not code)
    src = open(os.path.join(REF, "index.coffee")).read()
    vec = [dict(x=int(a), y=int(b), dx=float(c), dy=float(d)) for a, b, c, d in re.findall(
        r"\{ x: (\d+), y: (\d+), dx: (-?[\d.]+), dy: (-?[\d.]+) \}", src)]
    assert len(vec) == 24
    golden = {
        "source": "/root/reference/test/index.coffee:12-96",
        "options": dict(threshold=5, span=10),
        "cases": [
            dict(expect="s1_expected", target="s1_expected", revision="revision1", status="OK", height=279, width=280, vector=[]),
            dict(expect="s2_expected", target="s2_expected", revision="revision1", status="OK", height=117, width=180, vector=[]),
            dict(expect="s1_expected", target="s1_expected", revision="revision2", status="OK", height=279, width=280, vector=[]),
            dict(expect="s2_expected", target="s2_revision2", revision="revision2", status="SUSPICIOUS", height=117, width=180, vector=vec),
        ],
    }
    json.dump(golden, open(os.path.join(OUT, "reference_golden.json"), "w"), indent=1)

    flows = {}
    pairs = {
        "s1": (fx[("expected", "scenario1")], fx[("revision2", "scenario1")]),
        "s2r2": (fx[("expected", "scenario2")], fx[("revision2", "scenario2")]),
    }
    for name, (kind, W, H, seed, defect) in SYNTH.items():
        pairs[name] = synth.make_pair(kind, W, H, seed, defect)
    for name, (a, b) in pairs.items():
        for on, kw in OPTS.items():
            if name == "s1" and on not in ("default", "cfg4"):
                continue
            flows[f"{name}__{on}"] = cv2_flow(a, b, FlowParam(**kw))
    np.savez_compressed(os.path.join(OUT, "cv2_flows.npz"), **flows)
    json.dump(dict(opts=OPTS, synth=SYNTH, cv2=cv2.__version__, ipp=False),
              open(os.path.join(OUT, "cv2_flows_meta.json"), "w"), indent=1)
    print("wrote", OUT, {k: v.shape for k, v in flows.items()})


if __name__ == "__main__":
    main()
