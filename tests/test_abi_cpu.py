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
    tools; the two fixture PNGs are the reference's own test files); junk is an 'empty image'."""
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


def test_decode_jpeg_matches_cv2_goldens(tw):
    """tw_decode_gray, JPEG leg (row f-1): luma-only decode with libjpeg's ISLOW inverse DCT, bit-identical to
    cv2.imread(IMREAD_GRAYSCALE) on the reference's own progressive scenario1 fixture and on baseline / progressive files in
    every sampling (4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 / 4:1:1), with restart intervals, optimised tables, gray sources, 1x1
    (goldens: tools/make_golden_jpeg.py)."""
    import glob
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    img = tw.imread_gray(os.path.join(gold, "jpg", "fixture_s1_capture1.jpg"))
    assert img is not None and img.shape == (279, 280)
    assert np.array_equal(img, np.load(os.path.join(gold, "fixture_s1_expected.npy")))
    n = 0
    for jpg in sorted(glob.glob(os.path.join(gold, "jpg", "*.jpg"))):
        ref = jpg[:-4] + ".gray.npy"
        if os.path.exists(ref):
            got = tw.imread_gray(jpg)
            assert got is not None and np.array_equal(got, np.load(ref)), jpg
            n += 1
    assert n >= 11


def test_decode_jpeg_rejects_what_it_cannot_match(tw, tmp_path):
    """Truncated headers, arithmetic-coded / lossless frames and CMYK are 'empty images' (-> "Can't open <path>"), never guesses."""
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpg")
    data = open(os.path.join(gold, "base_420_q75.jpg"), "rb").read()
    sof = data.index(b"\xff\xc0")
    for name, blob in (("cut_header", data[:sof + 6]), ("arith", data[:sof] + b"\xff\xc9" + data[sof + 2:]),
                       ("lossless", data[:sof] + b"\xff\xc3" + data[sof + 2:]), ("no_scan", data[:data.index(b"\xff\xda")] + b"\xff\xd9"),
                       ("four_comp", data[:sof + 9] + b"\x04" + data[sof + 10:])):
        p = tmp_path / (name + ".jpg")
        p.write_bytes(blob)
        assert tw.imread_gray(str(p)) is None, name
    # a file cut inside the entropy-coded data still decodes (missing data reads as zero bits, like libjpeg) to the full size
    p = tmp_path / "cut_body.jpg"
    p.write_bytes(data[:len(data) - 200])
    got = tw.imread_gray(str(p))
    assert got is not None and got.shape == (61, 83)


def test_decode_jpeg_fuzz_vs_cv2(tw):
    """Live comparison with cv2 (when importable) over random sizes / qualities / coding variants."""
    cv2 = pytest.importorskip("cv2")
    import ctypes as C
    lib = tw.load()
    rng = np.random.default_rng(7)
    ss = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
          cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411]
    for it in range(60):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if it % 3 == 0:
            a = cv2.GaussianBlur(a, (5, 5), 1.5)
        gray = it % 7 == 0
        if gray:
            a = cv2.cvtColor(a, cv2.COLOR_BGR2GRAY)
        params = [cv2.IMWRITE_JPEG_QUALITY, int(rng.integers(5, 101)), cv2.IMWRITE_JPEG_PROGRESSIVE, int(rng.integers(0, 2)),
                  cv2.IMWRITE_JPEG_OPTIMIZE, int(rng.integers(0, 2)), cv2.IMWRITE_JPEG_RST_INTERVAL, int(rng.integers(0, 5))]
        if not gray:
            params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss[int(rng.integers(0, 5))]]
        ok, enc = cv2.imencode(".jpg", a, params)
        assert ok
        ref = cv2.imdecode(enc, cv2.IMREAD_GRAYSCALE)
        data = enc.tobytes()
        ww, hh = C.c_int(), C.c_int()
        assert lib.tw_decode_gray(data, len(data), None, 0, C.byref(ww), C.byref(hh)) == 0 and (hh.value, ww.value) == ref.shape
        out = np.empty(ref.shape, np.uint8)
        assert lib.tw_decode_gray(data, len(data), out.ctypes.data, out.size, C.byref(ww), C.byref(hh)) == 0
        assert np.array_equal(out, ref), (it, params)


def test_decode_fuzz_sanitizers(tmp_path):
    """Mutated PNG / JPEG goldens through tw_decode_gray under ASan + UBSan (tests/cpp/fuzz_decode.cpp): no out-of-bounds access,
    no undefined shift, nothing thrown across the C ABI."""
    import glob, shutil, subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fuzz_decode")
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17",
                           "-I" + os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "fuzz_decode.cpp"),
                           os.path.join(root, "tidal-wave_b200", "csrc", "tw_jpeg.cpp"), os.path.join(root, "tidal-wave_b200", "csrc", "tw_decode.cpp"),
                           "-lz", "-o", exe])
    files = sorted(glob.glob(os.path.join(root, "tests", "golden", "jpg", "*.jpg")) + glob.glob(os.path.join(root, "tests", "golden", "png", "*.png")))
    out = subprocess.run([exe] + files, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "runs" in out.stdout
