"""BASELINE.json configs[1..4] at FULL size on the B200, through the C ABI, against the C oracle (run with -m gpu).

  config 2  1920x1080, default options                      (S, T, S+defect)   both arithmetics
  config 3  3840x2160, pyrLevels 5, pyrIterations 5          (T, S)             both arithmetics -- the 4096-pitch kernel
            instantiations, the 16x / 32x pyramid levels with their 39- / 79-tap pre-blurs
  config 4  1280x2000, polyN 5, polySigma 1.1, winSize 15, flags 0 (box window)  faithful only (nothing may be relaxed there)
  config 5  >= 2000 pairs of the 1920x1080 pool through ONE tw_pool spanning every visible GPU, every answer compared with a
            single dense context's (the reference's Manager + Consumers: src/manager.cpp:55-59, src/consumer.cpp:18-24)

Bars: faithful arithmetic ("arithmetic" = 0) bit-identical to oracle/farneback_ref.c; the library default bit-identical to the
oracle's restatement of the same relaxation (twref_set_relax(144)) AND within the north-star tolerance (1e-2 px max, 1e-3 px
RMS) of the faithful oracle; status and vector sets identical everywhere (the assertions of
/root/reference/test/index.coffee:44-96, on the BASELINE workloads).
"""
import ctypes as C

import numpy as np
import pytest

from oracle.oracle import FlowParam

pytestmark = pytest.mark.gpu

TOL_MAX, TOL_RMS = 1e-2, 1e-3
CFG3 = dict(pyrLevels=5, pyrIterations=5)
CFG4 = dict(polyN=5, polySigma=1.1, winSize=15, flags=0)

CASES = [
    ("cfg2-S", ("S", 1920, 1080, 2, False), {}),
    ("cfg2-T", ("T", 1920, 1080, 1, False), {}),
    ("cfg2-Sdefect", ("S", 1920, 1080, 100, True), {}),
    ("cfg3-T", ("T", 3840, 2160, 4, False), CFG3),
    ("cfg3-S", ("S", 3840, 2160, 5, False), CFG3),
    ("cfg4-S", ("S", 1280, 2000, 3, False), CFG4),
    ("cfg4-T", ("T", 1280, 2000, 6, False), CFG4),
]


@pytest.fixture(scope="module")
def of4k(tw):
    o = tw.OpticalFlow(0, 3840, 2160, 1)
    yield o
    o.close()


def _vec(resp):
    return [(v["x"], v["y"], v["dx"], v["dy"]) for v in resp["vector"]]


@pytest.mark.parametrize("name,synth,kw", CASES, ids=[c[0] for c in CASES])
def test_fullsize_config(of4k, tw, oracle, name, synth, kw):
    kind, W, H, seed, defect = synth
    a, b = tw.synth.make_pair(kind, W, H, seed, defect)
    p = tw.OpticalFlowParameter(**kw)
    oracle.set_relax(0)
    ref = oracle.farneback(a, b, FlowParam(**kw))
    status, vec = oracle.sample(ref)
    if defect:
        assert status == "SUSPICIOUS"
    # ---- faithful arithmetic: bit-identical ----
    of4k.set_option("arithmetic", 0)
    rc, fx, fy, sec = of4k.calculateInternal(a, b, p)
    assert rc == 0, of4k.last_error()
    assert sec > 0
    assert np.array_equal(fx, ref[..., 0]) and np.array_equal(fy, ref[..., 1]), \
        f"{name} faithful: max |d| = {max(np.abs(fx - ref[..., 0]).max(), np.abs(fy - ref[..., 1]).max()):.3e}"
    for sparse in (0, 1):  # dense last iteration, and the dispatcher's classification-only last iteration
        of4k.set_option("sparse_last", sparse)
        resp = of4k.calculate(a, b, p)
        assert resp["status"] == status
        assert _vec(resp) == [(v[0], v[1], float(v[2]), float(v[3])) for v in vec], f"{name} faithful sparse_last={sparse}"
    of4k.set_option("sparse_last", 0)
    # ---- library default (relaxed where validated) ----
    of4k.set_option("arithmetic", 1)
    if of4k.arithmetic_in_effect(p) != "relaxed":
        assert kw.get("flags", 256) == 0  # config 4: the box window always runs the faithful kernels
        return
    oracle.set_relax(144)
    try:
        rel = oracle.farneback(a, b, FlowParam(**kw))
    finally:
        oracle.set_relax(0)
    rc, gx, gy, _ = of4k.calculateInternal(a, b, p)
    assert rc == 0, of4k.last_error()
    assert np.array_equal(gx, rel[..., 0]) and np.array_equal(gy, rel[..., 1]), f"{name} relaxed vs oracle(144)"
    d = np.maximum(np.abs(gx - ref[..., 0]), np.abs(gy - ref[..., 1]))
    rms = float(np.sqrt(((gx - ref[..., 0]) ** 2 + (gy - ref[..., 1]) ** 2).mean() / 2))
    assert d.max() <= TOL_MAX and rms <= TOL_RMS, f"{name} relaxed vs faithful oracle: max {d.max():.3e} rms {rms:.3e}"
    for sparse in (0, 1):
        of4k.set_option("sparse_last", sparse)
        resp = of4k.calculate(a, b, p)
        assert resp["status"] == status
        assert [(v["x"], v["y"]) for v in resp["vector"]] == [(v[0], v[1]) for v in vec]
        rstatus, rvec = oracle.sample(rel)
        assert _vec(resp) == [(v[0], v[1], float(v[2]), float(v[3])) for v in rvec]
    of4k.set_option("sparse_last", 0)


def test_config5_one_pool_all_gpus(tw):
    """configs[4] scaled to a test: 2048 pairs cycled from a pool of 32 distinct 1920x1080 pairs through ONE in-process
    dispatcher whose consumers cover every visible GPU (one consumer thread per GPU, as BASELINE's north_star states);
    mismatches against a single dense context must be zero and the Report must add up."""
    lib = tw.load()
    ngpu = lib.tw_device_count()
    W, H, POOL, N, B = 1920, 1080, 32, 2048, 16
    pairs = tw.synth.pool_pairs(POOL, W, H, seed0=100)
    of = tw.OpticalFlow(0, W, H, 8)  # dense last iteration (tw_create default)
    want = []
    for i in range(0, POOL, 8):
        want += [(r["status"], _vec(r)) for r in of.calculate_batch(pairs[i:i + 8])]
    of.close()
    assert any(s == "SUSPICIOUS" for s, _ in want) and any(s == "OK" for s, _ in want)
    pinned = []
    for a, b in pairs:
        pa = lib.tw_host_alloc(W * H); pb = lib.tw_host_alloc(W * H)
        C.memmove(pa, a.ctypes.data, W * H); C.memmove(pb, b.ctypes.data, W * H)
        pinned.append((pa, pb))
    devices = list(range(ngpu))
    p = tw.OpticalFlowParameter().c()
    err = C.create_string_buffer(256)
    pool = lib.tw_pool_create((C.c_int * len(devices))(*devices), len(devices), W, H, B, C.byref(p), 5.0, 10, 4096, err, 256)
    assert pool, err.value
    try:
        vec = (tw.tw_vector * 4096)(); res = tw.tw_result()
        ids = [lib.tw_pool_submit(pool, pinned[i % POOL][0], W, H, pinned[i % POOL][1], W, H) for i in range(N)]
        bad = 0
        for i, rid in enumerate(ids):
            rc = lib.tw_pool_wait(pool, rid, vec, 4096, C.byref(res))
            st, wv = want[i % POOL]
            got = [(vec[k].x, vec[k].y, vec[k].dx, vec[k].dy) for k in range(min(res.n_vectors, 4096))]
            if rc != 0 or tw.api.STATUS_NAMES[res.status] != st or got != wv or (res.width, res.height) != (W, H):
                bad += 1
        a = C.c_int(); b = C.c_int(); c = C.c_int()
        lib.tw_pool_report(pool, C.byref(a), C.byref(b), C.byref(c))
    finally:
        lib.tw_pool_destroy(pool)
        for pa, pb in pinned:
            lib.tw_host_free(pa); lib.tw_host_free(pb)
    assert bad == 0, f"{bad} of {N} answers differ from the single-context run ({ngpu} GPUs)"
    assert (a.value, b.value, c.value) == (N, N, 0)
