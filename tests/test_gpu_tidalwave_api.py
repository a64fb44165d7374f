"""File-level replay of /root/reference/test/index.coffee through the mirror of index.js (`create(targetDir, options)`): the
directory pairing + lifecycle and the PNG / JPEG decode (row f-1).  Both scenarios use the reference's own
fixture files: scenario2's PNGs and scenario1's progressive JPEG (byte-identical in all three revisions)."""
import os
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fixture_tree(tmp_path_factory, golden):
    root = tmp_path_factory.mktemp("fixture")
    for rev, s2 in (("expected", "fixture_s2_expected.png"), ("revision1", "fixture_s2_expected.png"), ("revision2", "fixture_s2_revision2.png")):
        os.makedirs(root / rev / "scenario2", exist_ok=True)
        shutil.copy(os.path.join(GOLD, "png", s2), root / rev / "scenario2" / "capture2.png")
        os.makedirs(root / rev / "scenario1", exist_ok=True)
        shutil.copy(os.path.join(GOLD, "jpg", "fixture_s1_capture1.jpg"), root / rev / "scenario1" / "capture1.jpg")
    return root


def _collect(t, tw):
    data, errors, fin = [], [], []
    t.on("data", data.append).on("error", errors.append).on("finish", fin.append)
    tw.run(t)
    return data, errors, fin


@pytest.mark.parametrize("rev", ["revision1", "revision2"])
def test_expect_dir(tw, golden, fixture_tree, rev):
    """'should report nothing on revision1' / 'should report something on revision2' (index.coffee:12-96)."""
    t = tw.create(str(fixture_tree / rev), {"expectDir": str(fixture_tree / "expected")})
    data, errors, fin = _collect(t, tw)
    assert errors == [] and fin == [{"request": 2, "data": 2, "error": 0}]
    for d in data:
        d.pop("time")
        if "capture1" in d["target_image"]:
            assert d == {"status": "OK", "span": 10, "threshold": 5, "height": 279, "width": 280, "vector": [],
                         "expect_image": str(fixture_tree / "expected/scenario1/capture1.jpg"),
                         "target_image": str(fixture_tree / rev / "scenario1/capture1.jpg")}
        else:
            want = [c for c in golden["cases"] if c["revision"] == rev and c["width"] == 180][0]
            assert (d["status"], d["height"], d["width"], d["span"], d["threshold"]) == (want["status"], 117, 180, 10, 5)
            assert d["expect_image"] == str(fixture_tree / "expected/scenario2/capture2.png")
            assert [(v["x"], v["y"]) for v in d["vector"]] == [(g["x"], g["y"]) for g in want["vector"]]
            for v, g in zip(d["vector"], want["vector"]):
                # library default arithmetic (relaxed): 1.9e-3 px from the reference's goldens on the -88 px vector (cv2 itself:
                # 3.3e-4; north-star bar 1e-2); the faithful arithmetic is pinned at 1e-3 in tests/test_gpu_parity.py
                assert abs(v["dx"] - g["dx"]) < 3e-3 and abs(v["dy"] - g["dy"]) < 3e-3


def test_missing_target_dir(tw, fixture_tree):
    """'should never report on __NOT_EXISTS__' (index.coffee:98-104)."""
    t = tw.create(str(fixture_tree / "__NOT_EXISTS__"), {"expectDir": str(fixture_tree / "expected")})
    data, errors, fin = _collect(t, tw)
    assert data == [] and fin == [{"request": 0, "data": 0, "error": 0}]


def test_get_expected_path_option(tw, fixture_tree):
    """'should start with passing getExpectedPath option' (index.coffee:106-117)."""
    t = tw.create(str(fixture_tree / "revision2"), {"getExpectedPath": lambda short: str(fixture_tree / "revision1" / short)})
    data, errors, fin = _collect(t, tw)
    assert len(data) == 2 and fin == [{"request": 2, "data": 2, "error": 0}]


def test_option_parsing_and_errors(tw, fixture_tree, tmp_path):
    """Wrong-typed options fall back to the defaults (src/broker.cpp:190-209); undecodable / missing files are errors."""
    t = tw.TidalWave({"threshold": "5", "span": 7.5, "polyN": 5, "polySigma": 1.1})
    assert (t.threshold, t.span, t.param.polyN, t.param.polySigma, t.param.winSize, t.param.flags) == (5.0, 10, 5, 1.1, 30, 256)
    with pytest.raises(ValueError):
        tw.create(str(fixture_tree), {})
    tgt = tmp_path / "tgt"; os.makedirs(tgt / "a")
    shutil.copy(os.path.join(GOLD, "png", "fixture_s2_revision2.png"), tgt / "a" / "x.png")
    (tgt / "a" / "y.jpg").write_bytes(b"\xff\xd8\xff\xe0 not decodable here")
    t = tw.create(str(tgt), {"expectDir": str(fixture_tree / "expected" / "scenario2" / ".."), "span": 20})
    data, errors, fin = _collect(t, tw)
    assert data == [] and fin == [{"request": 2, "data": 0, "error": 2}]
    assert all(e["status"] == "ERROR" and e["reason"].startswith("Can't open ") for e in errors)
