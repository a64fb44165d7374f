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


def _png_bytes(samples, ctype, depth, interlace, rng, palette=None, trns=None, filters=None):
    """A PNG file of the H x W x C integer samples (< 2**depth), written here so that every colour type / bit depth / Adam7 /
    row filter the format has can be put in front of the decoder (cv2.imwrite only writes a few of them)."""
    import struct
    import zlib
    H, W, C = samples.shape
    bits = C * depth
    bpp = max(1, bits // 8)

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    def pack_rows(img):
        h, w, _ = img.shape
        flat = img.reshape(h, w * C)
        if depth == 16:
            return np.stack([flat >> 8, flat & 255], -1).reshape(h, -1).astype(np.uint8)
        if depth == 8:
            return flat.astype(np.uint8)
        per = 8 // depth
        pad = (-flat.shape[1]) % per
        flat = np.pad(flat, ((0, 0), (0, pad)))
        out = np.zeros((h, flat.shape[1] // per), np.int64)
        for k in range(per):
            out |= flat[:, k::per] << ((per - 1 - k) * depth)
        return out.astype(np.uint8)

    def filt(rows):
        out = bytearray()
        prev = np.zeros(rows.shape[1], np.int64)
        for y in range(rows.shape[0]):
            cur = rows[y].astype(np.int64)
            f = int(rng.integers(0, 5)) if filters is None else filters
            a = np.concatenate([np.zeros(bpp, np.int64), cur[:-bpp]]) if cur.size > bpp else np.zeros_like(cur)
            if cur.size <= bpp:
                a = np.zeros_like(cur)
            c = np.concatenate([np.zeros(bpp, np.int64), prev[:-bpp]]) if cur.size > bpp else np.zeros_like(cur)
            if f == 0:
                pred = 0
            elif f == 1:
                pred = a
            elif f == 2:
                pred = prev
            elif f == 3:
                pred = (a + prev) >> 1
            else:
                pp = a + prev - c
                pa, pb, pc = np.abs(pp - a), np.abs(pp - prev), np.abs(pp - c)
                pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
            out.append(f)
            out += ((cur - pred) & 255).astype(np.uint8).tobytes()
            prev = cur
        return bytes(out)

    if interlace:
        passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]
    else:
        passes = [(0, 0, 1, 1)]
    raw = b""
    for x0, y0, dx, dy in passes:
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] and sub.shape[1]:
            raw += filt(pack_rows(sub))
    comp = zlib.compress(raw, 6)
    cut = len(comp) // 2
    body = chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None:
        body += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    if trns is not None:
        body += chunk(b"tRNS", trns)
    body += chunk(b"IDAT", comp[:cut]) + chunk(b"IDAT", comp[cut:]) + chunk(b"IEND", b"")
    return b"\x89PNG\r\n\x1a\n" + body


def _png_expected_gray(samples, ctype, depth, palette):
    """What cv2.imread(IMREAD_GRAYSCALE) makes of those samples (tw_decode.cpp's header comment), restated in NumPy."""
    s = samples.astype(np.int64)
    if ctype == 3:
        rgb = palette.astype(np.int64)[s[..., 0]]
        return ((9797 * rgb[..., 0] + 19234 * rgb[..., 1] + 3737 * rgb[..., 2]) >> 15).astype(np.uint8)
    if ctype in (0, 4):
        g = s[..., 0]
        return (g >> 8 if depth == 16 else g * (255 // ((1 << depth) - 1))).astype(np.uint8)
    r, g, b = s[..., 0], s[..., 1], s[..., 2]
    if depth == 16:
        return (((9797 * r + 19234 * g + 3737 * b + 16384) >> 15) >> 8).astype(np.uint8)
    return ((9797 * r + 19234 * g + 3737 * b) >> 15).astype(np.uint8)


def test_decode_png_every_type_depth_interlace(tw):
    """Every (colour type, bit depth) pair of the PNG format, Adam7-interlaced or not, random row filters, ragged sizes (1-pixel and
    sub-8 widths leave Adam7 passes empty): equal to the NumPy restatement of OpenCV's conversion, and to cv2 itself when importable."""
    import ctypes as C
    try:
        import cv2
    except ImportError:
        cv2 = None
    lib = tw.load()
    rng = np.random.default_rng(23)
    combos = [(0, d) for d in (1, 2, 4, 8, 16)] + [(2, 8), (2, 16)] + [(3, d) for d in (1, 2, 4, 8)] + [(4, 8), (4, 16), (6, 8), (6, 16)]
    sizes = [(1, 1), (1, 9), (7, 3), (8, 8), (13, 21), (33, 50), (64, 37), (4, 333)]  # the last: rows long enough for the SIMD Sub / gray paths
    nchan = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}
    checked = 0
    for ctype, depth in combos:
        for interlace in (0, 1):
            for (h, w) in sizes:
                samples = rng.integers(0, 1 << depth, (h, w, nchan[ctype]))
                if w > 100:  # screenshot-like structure: flat runs and repeated rows (the degenerate Paeth groups of the SIMD filter path)
                    samples[:, w // 4: w // 2] = samples[:, w // 4: w // 4 + 1]
                    samples[2:] = samples[1]
                    samples[3, w // 3] += 1
                    samples %= 1 << depth
                palette = rng.integers(0, 256, (1 << depth, 3)) if ctype == 3 else None
                trns = None
                if rng.integers(0, 2):  # transparency information must not change the gray output
                    if ctype == 3:
                        trns = rng.integers(0, 256, int(rng.integers(1, (1 << depth) + 1)), dtype=np.uint8).tobytes()
                    elif ctype == 0:
                        trns = int(rng.integers(0, 1 << depth)).to_bytes(2, "big")
                    elif ctype == 2:
                        trns = b"".join(int(v).to_bytes(2, "big") for v in rng.integers(0, 1 << depth, 3))
                data = _png_bytes(samples, ctype, depth, interlace, rng, palette, trns)
                want = _png_expected_gray(samples, ctype, depth, palette)
                ww, hh = C.c_int(), C.c_int()
                out = np.empty((h, w), np.uint8)
                rc = lib.tw_decode_gray(data, len(data), out.ctypes.data, out.size, C.byref(ww), C.byref(hh))
                assert rc == 0 and (hh.value, ww.value) == (h, w), (ctype, depth, interlace, h, w, rc)
                assert np.array_equal(out, want), (ctype, depth, interlace, h, w)
                if cv2 is not None:
                    ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
                    assert ref is not None and np.array_equal(out, ref), ("cv2", ctype, depth, interlace, h, w)
                checked += 1
    assert checked == len(combos) * 2 * len(sizes)


def test_decode_png_rejects_damaged_streams(tw):
    """A wrong CRC on a critical chunk, a truncated IDAT stream and an unknown filter type fail like libpng's read does (the
    callers turn that into "Can't open <path>"); data after the last scanline is tolerated (libpng: a warning)."""
    import ctypes as C
    import struct
    import zlib
    lib = tw.load()
    rng = np.random.default_rng(3)
    samples = rng.integers(0, 256, (9, 11, 3))
    good = _png_bytes(samples, 2, 8, 0, rng)
    out = np.empty((9, 11), np.uint8)
    ww, hh = C.c_int(), C.c_int()

    def rc(data):
        return lib.tw_decode_gray(data, len(data), out.ctypes.data, out.size, C.byref(ww), C.byref(hh))
    assert rc(good) == 0
    i = good.index(b"IDAT")
    bad_crc = bytearray(good)
    bad_crc[i + 6] ^= 0x40
    assert rc(bytes(bad_crc)) != 0
    # rebuild the file around a modified raw stream
    def rebuild(raw):
        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
        return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 11, 9, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw))
                + chunk(b"IEND", b""))
    raw = b"".join(b"\x00" + samples[y].astype(np.uint8).tobytes() for y in range(9))
    assert rc(rebuild(raw)) == 0
    assert rc(rebuild(raw[:-5])) != 0                 # not enough image data
    assert rc(rebuild(raw + b"\x00" * 40)) == 0       # too much: ignored
    assert rc(rebuild(b"\x07" + raw[1:])) != 0        # filter type 7 does not exist


def test_inflate_matches_zlib(tmp_path):
    """tw_inflate.h (the PNG leg's own zlib-stream decoder) against zlib under ASan + UBSan: tests/cpp/inflate_diff.cpp."""
    import shutil, subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "inflate_diff")
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17",
                           os.path.join(root, "tests", "cpp", "inflate_diff.cpp"), "-lz", "-o", exe])
    out = subprocess.run([exe, "5", "500"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-2000:])
    assert "cases 500 bad 0" in out.stdout
