"""The reference's own integration test (/root/reference/test/index.coffee:12-117) replayed through the C++ mirror of its
Manager / Consumer / OpticalFlow classes (tests/cpp/tidalwave_host.hpp) on top of the C ABI."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host")


def _pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def _run(args):
    out = subprocess.run([EXE] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


@pytest.fixture(scope="module")
def files(tmp_path_factory, golden):
    d = tmp_path_factory.mktemp("pgm")
    paths = {}
    for name, img in golden["imgs"].items():
        paths[name] = str(d / f"{name}.pgm")
        _pgm(paths[name], img)
    return paths


def test_host_built():
    assert os.path.exists(EXE), "run __graft_entry__.build()"


@pytest.mark.parametrize("threads", [1, 2, 4])
def test_reference_integration_cases(files, golden, threads):
    """'should report nothing on revision1' + 'should report something on revision2' (index.coffee:12-96)."""
    args = [threads]
    for case in golden["cases"]:
        args += [files[case["expect"]], files[case["target"]]]
    ev = _run(args)
    data = [e for e in ev if e["event"] == "data"]
    fin = [e for e in ev if e["event"] == "finish"]
    assert fin == [{"event": "finish", "request": 4, "data": 4, "error": 0}]
    by_pair = {(e["expect_image"], e["target_image"]): e for e in data}
    for case in golden["cases"]:
        e = by_pair[(files[case["expect"]], files[case["target"]])]
        assert e["status"] == case["status"] and (e["height"], e["width"], e["span"], e["threshold"]) == (case["height"], case["width"], 10, 5)
        assert [(v["x"], v["y"]) for v in e["vector"]] == [(g["x"], g["y"]) for g in case["vector"]]
        for v, g in zip(e["vector"], case["vector"]):
            # library default arithmetic (relaxed): 1.9e-3 px from the reference's goldens on the -88 px vector (cv2 itself:
            # 3.3e-4; north-star bar 1e-2); the faithful arithmetic is pinned at 1e-3 in tests/test_gpu_parity.py
            assert abs(v["dx"] - g["dx"]) < 3e-3 and abs(v["dy"] - g["dy"]) < 3e-3


def test_errors_and_size_rule(files, golden, tmp_path):
    """Error values of src/opticalflow.cpp:26-61 and the +-5 px rule through Manager/Consumer; Report counts them."""
    a = golden["imgs"]["s2_expected"]
    big = str(tmp_path / "big.pgm"); _pgm(big, np.zeros((a.shape[0] + 6, a.shape[1]), np.uint8))
    near = str(tmp_path / "near.pgm"); _pgm(near, np.pad(a, ((0, 3), (0, 2)), mode="edge"))
    junk = str(tmp_path / "junk.pgm"); open(junk, "wb").write(b"not an image")
    ev = _run([2, files["s2_expected"], str(tmp_path / "missing.pgm"), files["s2_expected"], big, junk, files["s2_expected"],
               files["s2_expected"], near])
    errs = sorted(e["reason"] for e in ev if e["event"] == "error")
    assert errs == sorted(["Can't open " + str(tmp_path / "missing.pgm"), "Don't match image size", "Can't open " + junk])
    data = [e for e in ev if e["event"] == "data"]
    assert len(data) == 1 and data[0]["target_image"] == near and (data[0]["height"], data[0]["width"]) == a.shape
    assert [e for e in ev if e["event"] == "finish"] == [{"event": "finish", "request": 4, "data": 1, "error": 3}]


def test_nothing_requested(files):
    """'should never report on __NOT_EXISTS__' (index.coffee:98-104): no requests -> report all zeros."""
    ev = _run([2])
    assert ev == [{"event": "finish", "request": 0, "data": 0, "error": 0}]
