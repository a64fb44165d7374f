import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# option sets of tests/golden/cv2_flows_meta.json (SURVEY App. A header) -- default = BASELINE config 1/2,
# cfg4 = BASELINE config 4, cfg3 = BASELINE config 3's options
OPTS = {
    "default": dict(),
    "cfg4": dict(polyN=5, polySigma=1.1, winSize=15, flags=0),
    "cfg3": dict(pyrLevels=5, pyrIterations=5),
    "box31": dict(flags=0),
    "g15n5": dict(polyN=5, polySigma=1.1, winSize=15),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    g = json.load(open(os.path.join(GOLDEN, "reference_golden.json")))
    imgs = {n: np.load(os.path.join(GOLDEN, f"fixture_{n}.npy")) for n in ("s1_expected", "s2_expected", "s2_revision2")}
    flows = np.load(os.path.join(GOLDEN, "cv2_flows.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "cv2_flows_meta.json")))
    return dict(cases=g["cases"], options=g["options"], imgs=imgs, flows=flows, meta=meta)


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import RefOracle
    return RefOracle()


@pytest.fixture(scope="session")
def tw():
    import tidalwave_b200
    return tidalwave_b200


def golden_pair(golden, tw, name):
    """name in {'s1','s2r2'} (reference fixtures) or a key of meta['synth'] (seeded synthetic)."""
    if name == "s1":
        return golden["imgs"]["s1_expected"], golden["imgs"]["s1_expected"]
    if name == "s2r2":
        return golden["imgs"]["s2_expected"], golden["imgs"]["s2_revision2"]
    kind, W, H, seed, defect = golden["meta"]["synth"][name]
    return tw.synth.make_pair(kind, W, H, seed, defect)
