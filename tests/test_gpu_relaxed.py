"""GPU tests of the library DEFAULT arithmetic (tw_set_option "arithmetic" = 1: relaxed where validated) and of the
CUDA-graph replay of the launch sequence.  Everything goes through the C ABI.

The relaxed kernels are checked twice:
  * bit for bit against the SAME relaxation restated in the oracle (oracle/farneback_ref.c, twref_set_relax(144)) --
    so the kernels compute exactly the documented arithmetic, and
  * within the north-star tolerance (BASELINE.json: 1e-2 px max-abs, 1e-3 px RMS, identical classification) against the
    FAITHFUL oracle and the committed cv2 flows.
Outside the validated option family (box window, smaller windows, polyN != 7) the default must stay bit-identical to the
faithful oracle."""
import numpy as np
import pytest

from conftest import OPTS, golden_pair
from oracle.oracle import FlowParam, sample_numpy

pytestmark = pytest.mark.gpu

TOL_MAX, TOL_RMS = 1e-2, 1e-3
RELAX_BITS = 144  # direct-form fmaf window taps (128) + mixed double/float horizontal poly-exp pass (16)
# The reference's scenario2 fixture iterated 5 times per scale is chaotic for EVERY implementation: cv2 4.13.0 itself moves by
# 0.24 px between cv2.setUseOptimized(True) and (False) on it (0.014 px at 4 iterations, 7e-4 px at the default 3), so no
# build of the reference's library pins it to 1e-2 px.  There the relaxed kernels are held to bit-equality with the restated
# relaxation, to identical classification, and to the same order of deviation as cv2's own two code paths.
CHAOTIC = {("s2r2", "cfg3"): (0.5, 1e-2)}


@pytest.fixture(scope="module")
def of(tw):
    tw.set_default_arithmetic(True)
    o = tw.OpticalFlow(0, 1920, 1080, 2)
    yield o
    o.close()


def _dev(fx, fy, ref):
    d = np.maximum(np.abs(fx - ref[..., 0]), np.abs(fy - ref[..., 1]))
    rms = float(np.sqrt(((fx - ref[..., 0]) ** 2 + (fy - ref[..., 1]) ** 2).mean() / 2))
    return float(d.max()), rms


@pytest.mark.parametrize("name", ["s1", "s2r2", "S256", "T256", "Sdef"])
@pytest.mark.parametrize("opt", ["default", "cfg3"])
def test_relaxed_default(of, tw, oracle, golden, name, opt):
    a, b = golden_pair(golden, tw, name)
    p = tw.OpticalFlowParameter(**OPTS[opt])
    assert of.arithmetic_in_effect(p) == "relaxed"
    rc, fx, fy, _ = of.calculateInternal(a, b, p)
    assert rc == 0, of.last_error()
    # (1) exactly the documented relaxation
    oracle.set_relax(RELAX_BITS)
    try:
        rel = oracle.farneback(a, b, FlowParam(**OPTS[opt]))
    finally:
        oracle.set_relax(0)
    assert np.array_equal(fx, rel[..., 0]) and np.array_equal(fy, rel[..., 1]), f"{name}/{opt}: not the documented relaxation"
    # (2) within tolerance of the faithful oracle and of cv2, classification identical
    ref = oracle.farneback(a, b, FlowParam(**OPTS[opt]))
    tol_max, tol_rms = CHAOTIC.get((name, opt), (TOL_MAX, TOL_RMS))
    mx, rms = _dev(fx, fy, ref)
    assert mx <= tol_max and rms <= tol_rms, f"{name}/{opt} vs faithful oracle: max {mx:.3e} rms {rms:.3e}"
    key = f"{name}__{opt}"
    if key in golden["flows"]:
        mx, rms = _dev(fx, fy, golden["flows"][key])
        assert mx <= tol_max and rms <= tol_rms, f"{name}/{opt} vs cv2: max {mx:.3e} rms {rms:.3e}"
    resp = of.calculate(a, b, p)
    status, vec = oracle.sample(ref)
    assert resp["status"] == status
    assert [(v["x"], v["y"]) for v in resp["vector"]] == [(v[0], v[1]) for v in vec]


def test_relaxed_reference_goldens(of, golden):
    """/root/reference/test/index.coffee:12-96 under the default arithmetic."""
    thr, span = golden["options"]["threshold"], golden["options"]["span"]
    for case in golden["cases"]:
        a, b = golden["imgs"][case["expect"]], golden["imgs"][case["target"]]
        resp = of.calculate(a, b, threshold=thr, span=span)
        assert resp["status"] == case["status"]
        assert [(v["x"], v["y"]) for v in resp["vector"]] == [(g["x"], g["y"]) for g in case["vector"]]
        for v, g in zip(resp["vector"], case["vector"]):
            assert abs(v["dx"] - g["dx"]) < 3e-3 and abs(v["dy"] - g["dy"]) < 3e-3


@pytest.mark.parametrize("kw", [OPTS["cfg4"], OPTS["box31"], OPTS["g15n5"], dict(winSize=20), dict(polyN=5, polySigma=1.1)])
def test_relaxation_scope(of, tw, oracle, kw):
    """Option sets outside the validated family keep the faithful kernels: bit-identical to the faithful oracle."""
    a, b = tw.synth.make_pair("S", 300, 220, 31, defect=True)
    p = tw.OpticalFlowParameter(**kw)
    assert of.arithmetic_in_effect(p) == "faithful"
    rc, fx, fy, _ = of.calculateInternal(a, b, p)
    assert rc == 0, of.last_error()
    ref = oracle.farneback(a, b, FlowParam(**kw))
    assert np.array_equal(fx, ref[..., 0]) and np.array_equal(fy, ref[..., 1]), kw


def test_arithmetic_option_switches(tw, oracle):
    a, b = tw.synth.make_pair("S", 480, 300, 33, defect=True)
    ref = oracle.farneback(a, b, FlowParam())
    oracle.set_relax(RELAX_BITS)
    rel = oracle.farneback(a, b, FlowParam())
    oracle.set_relax(0)
    o = tw.OpticalFlow(0, 480, 300, 1)
    for mode, want in ((1, rel), (0, ref), (1, rel)):
        o.set_option("arithmetic", mode)
        for _ in range(3):  # eager, capture, replay
            rc, fx, fy, _ = o.calculateInternal(a, b)
            assert rc == 0 and np.array_equal(fx, want[..., 0]) and np.array_equal(fy, want[..., 1]), mode
    o.close()


def test_relaxed_full_size_1080p(tw, oracle):
    """BASELINE configs[1] at full size (screenshot-like pair with a defect): relaxed default vs both oracles."""
    a, b = tw.synth.make_pair("S", 1920, 1080, 100, True)
    o = tw.OpticalFlow(0, 1920, 1080, 1)
    rc, fx, fy, _ = o.calculateInternal(a, b)
    assert rc == 0
    ref = oracle.farneback(a, b, FlowParam())
    mx, rms = _dev(fx, fy, ref)
    assert mx <= TOL_MAX and rms <= TOL_RMS, f"max {mx:.3e} rms {rms:.3e}"
    oracle.set_relax(RELAX_BITS)
    rel = oracle.farneback(a, b, FlowParam())
    oracle.set_relax(0)
    assert np.array_equal(fx, rel[..., 0]) and np.array_equal(fy, rel[..., 1])
    resp = o.calculate(a, b)
    status, vec = oracle.sample(ref)
    assert resp["status"] == status == "SUSPICIOUS"
    assert [(v["x"], v["y"]) for v in resp["vector"]] == [(v[0], v[1]) for v in vec]
    o.close()


@pytest.mark.parametrize("arith", [0, 1])
def test_graph_replay_equals_eager(tw, arith):
    """Runs 2+ of an unchanged (size, batch, options) key replay a captured CUDA graph: same results, same launch count."""
    pairs = [tw.synth.make_pair("S", 640, 360, 50 + i, defect=(i % 2 == 0)) for i in range(3)]
    eager = tw.OpticalFlow(0, 640, 360, 3)
    eager.set_option("graph", 0)
    eager.set_option("arithmetic", arith)
    want = eager.calculate_batch(pairs)
    wfx, wfy = eager.batch_flow(1, 640, 360)
    n0 = eager.launch_count(); eager.calculate_batch(pairs); per_run = eager.launch_count() - n0
    eager.close()
    g = tw.OpticalFlow(0, 640, 360, 3)
    g.set_option("arithmetic", arith)
    for rep in range(4):
        n0 = g.launch_count()
        got = g.calculate_batch(pairs)
        assert g.launch_count() - n0 == per_run, rep
        for w_, r in zip(want, got):
            assert w_["status"] == r["status"] and w_["vector"] == r["vector"], rep
        fx, fy = g.batch_flow(1, 640, 360)
        assert np.array_equal(fx, wfx) and np.array_equal(fy, wfy), rep
    # a different threshold / span is a different key (eager again, then its own graph)
    r = g.calculate_batch(pairs, threshold=2.0, span=7)
    r2 = g.calculate_batch(pairs, threshold=2.0, span=7)
    r3 = g.calculate_batch(pairs, threshold=2.0, span=7)
    assert [x["vector"] for x in r] == [x["vector"] for x in r2] == [x["vector"] for x in r3]
    # other sizes rebuild the plan and drop the graph
    small = [tw.synth.make_pair("T", 320, 200, 9)]
    for _ in range(3):
        assert g.calculate_batch(small)[0]["status"] == "OK"
    got = g.calculate_batch(pairs)
    assert [x["vector"] for x in got] == [x["vector"] for x in want]
    g.close()


@pytest.mark.gpu
def test_polyexp_tma_tiles_bit_identical(tw, oracle):
    """"polyexp_tma" = 1 stages the interior input tiles of the relaxed polynomial expansion with a TMA tensor-map copy; border tiles keep
    the per-thread loads.  Same values in, same arithmetic: bit-identical to oracle(144) on a frame with interior and border tiles."""
    a, b = tw.synth.make_pair("S", 700, 300, 41, defect=True)
    oracle.set_relax(RELAX_BITS)
    rel = oracle.farneback(a, b, FlowParam())
    oracle.set_relax(0)
    for tma in (1, 0):
        o = tw.OpticalFlow(0, 700, 300, 1)
        o.set_option("polyexp_tma", tma)
        for _ in range(2):
            rc, fx, fy, _ = o.calculateInternal(a, b)
            assert rc == 0 and np.array_equal(fx, rel[..., 0]) and np.array_equal(fy, rel[..., 1]), tma
        o.close()
