"""CPU tests of the boundary: the C-ABI library loads, exports every declared symbol, struct layouts match the reference's,
and compute fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os

import numpy as np
import pytest


def test_library_built_and_exports_all_symbols(tw):
    if not os.path.exists(tw.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = tw.load()
    syms = tw.declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tidalwave_b200.h but not exported"
    assert b"sm_100a" in lib.tw_version()


def test_struct_layouts(tw):
    # OpticalFlowParameter (src/opticalflow.h:28-36): double,int,int,int,int,double,int -> 40 bytes on LP64
    assert C.sizeof(tw.tw_flow_param) == 40
    assert tw.tw_flow_param.polySigma.offset == 24
    # Vector (src/message_queue.h:20-25): int,int,double,double -> 24 bytes
    assert C.sizeof(tw.tw_vector) == 24
    assert C.sizeof(tw.tw_result) == 24 + 128


def test_defaults_match_broker(tw):
    lib = tw.load()
    p = tw.tw_flow_param()
    lib.tw_default_param(C.byref(p))
    # src/broker.cpp:106-117
    assert (p.pyrScale, p.pyrLevels, p.winSize, p.pyrIterations, p.polyN, p.polySigma, p.flags) == (0.5, 3, 30, 3, 7, 1.5, 256)
    d = tw.OpticalFlowParameter()
    assert (d.pyrScale, d.pyrLevels, d.winSize, d.pyrIterations, d.polyN, d.polySigma, d.flags) == (0.5, 3, 30, 3, 7, 1.5, 256)


def test_no_cpu_fallback(tw):
    """Without a CUDA device the operator must refuse to exist rather than compute on the host."""
    lib = tw.load()
    if lib.tw_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        tw.OpticalFlow(0)
    with pytest.raises(RuntimeError):
        tw.Pool([0])


def test_product_does_not_link_oracle(tw):
    """The product library must not reference the oracle (prompt section 3)."""
    data = open(tw.LIB_PATH, "rb").read()
    assert b"twref_" not in data and b"libtwref" not in data
    api = open(os.path.join(os.path.dirname(tw.LIB_PATH), "api.py")).read()
    assert "oracle" not in api.replace("no CPU fallback", "")


def test_synth_is_deterministic(tw):
    a1, b1 = tw.synth.make_pair("S", 160, 120, 5, defect=True)
    a2, b2 = tw.synth.make_pair("S", 160, 120, 5, defect=True)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2) and a1.dtype == np.uint8 and not np.array_equal(a1, b1)


def test_decode_gray_matches_cv2_fixtures(tw):
    """tw_decode_gray (row f-1): PNG colour types 0/2/3/6 -> gray bit-identical to cv2.imread(IMREAD_GRAYSCALE) (goldens made by
    tools; the two fixture PNGs are the reference's own test files); JPEG and junk are 'empty images'."""
    import glob
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    n = 0
    for png in sorted(glob.glob(os.path.join(gold, "png", "*.png"))):
        ref = png[:-4] + ".gray.npy"
        img = tw.imread_gray(png)
        assert img is not None and img.dtype == np.uint8
        if os.path.exists(ref):
            assert np.array_equal(img, np.load(ref)), png
            n += 1
    assert n >= 5
    assert np.array_equal(tw.imread_gray(os.path.join(gold, "png", "fixture_s2_expected.png")), np.load(os.path.join(gold, "fixture_s2_expected.npy")))
    assert np.array_equal(tw.imread_gray(os.path.join(gold, "png", "fixture_s2_revision2.png")), np.load(os.path.join(gold, "fixture_s2_revision2.npy")))
    assert tw.imread_gray("/nonexistent/file.png") is None


def test_decode_pgm_roundtrip(tw, tmp_path):
    img = (np.arange(37 * 53) % 251).astype(np.uint8).reshape(37, 53)
    p = tmp_path / "a.pgm"
    p.write_bytes(b"P5\n# comment\n53 37\n255\n" + img.tobytes())
    assert np.array_equal(tw.imread_gray(str(p)), img)
    (tmp_path / "b.jpg").write_bytes(b"\xff\xd8\xff\xe0junk")
    assert tw.imread_gray(str(tmp_path / "b.jpg")) is None
