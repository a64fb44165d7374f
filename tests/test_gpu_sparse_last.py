"""GPU tests of the classification-only last iteration ("sparse_last", kernel K5s): blur + solve of the finest scale's last
iteration evaluated only at the positions /root/reference/src/consumer.cpp:60-77 samples.  The vectors (positions AND the
float dx, dy) and the status must be bit-identical to the dense path and to the oracle in both arithmetics; the dispatcher
(tw_pool) uses it by default."""
import numpy as np
import pytest

from conftest import OPTS
from oracle.oracle import FlowParam

pytestmark = pytest.mark.gpu


def _vec(resp):
    return [(v["x"], v["y"], v["dx"], v["dy"]) for v in resp["vector"]]


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("opt,size,span,thr", [("default", (333, 217), 10, 0.3), ("default", (640, 360), 7, 0.05), ("default", (226, 130), 4, 0.0),
                                               ("default", (1921, 131), 10, 0.0), ("g15n5", (300, 200), 10, 0.2), ("cfg3", (512, 320), 13, 0.1)])
def test_sparse_last_equals_dense(tw, arith, opt, size, span, thr):
    w, h = size
    pairs = [tw.synth.make_pair("S" if i % 2 == 0 else "T", w, h, 70 + i, defect=(i == 0)) for i in range(3)]
    p = tw.OpticalFlowParameter(**OPTS[opt])
    o = tw.OpticalFlow(0, w, h, 3)
    o.set_option("arithmetic", arith)
    dense = o.calculate_batch(pairs, p, threshold=thr, span=span)
    fx, fy = o.batch_flow(0, w, h)
    o.set_option("sparse_last", 1)
    for rep in range(3):  # eager, graph capture, graph replay
        sparse = o.calculate_batch(pairs, p, threshold=thr, span=span)
        for d, s in zip(dense, sparse):
            assert d["status"] == s["status"] and (d["width"], d["height"]) == (s["width"], s["height"])
            assert _vec(d) == _vec(s), (rep, len(d["vector"]), len(s["vector"]))
        with pytest.raises(RuntimeError):
            o.batch_flow(0, w, h)  # no dense field after a sparse run
    assert sum(len(d["vector"]) for d in dense) > 0  # the thresholds above make the comparison non-trivial
    # the vectors are the dense flow at the sampled positions
    for v in dense[0]["vector"]:
        assert v["dx"] == float(fx[v["y"], v["x"]]) and v["dy"] == float(fy[v["y"], v["x"]])
    # back to dense: the field is there again
    o.set_option("sparse_last", 0)
    again = o.calculate_batch(pairs, p, threshold=thr, span=span)
    assert [_vec(d) for d in again] == [_vec(d) for d in dense]
    fx2, fy2 = o.batch_flow(0, w, h)
    assert np.array_equal(fx, fx2) and np.array_equal(fy, fy2)
    o.close()


def test_sparse_last_vs_oracle(tw, oracle):
    a, b = tw.synth.make_pair("S", 480, 300, 33, defect=True)
    o = tw.OpticalFlow(0, 480, 300, 1)
    o.set_option("arithmetic", 0)
    o.set_option("sparse_last", 1)
    ref = oracle.farneback(a, b, FlowParam())
    for thr, span in ((5.0, 10), (0.2, 10), (0.0, 6)):
        status, vec = oracle.sample(ref, span=span, threshold=thr)
        resp = o.calculate_batch([(a, b)], threshold=thr, span=span)[0]
        assert resp["status"] == status
        assert [(v["x"], v["y"], v["dx"], v["dy"]) for v in resp["vector"]] == [tuple(v) for v in vec]
    o.close()


def test_box_window_and_other_radii_stay_dense(tw):
    """Option sets the sparse kernel does not cover run the dense path unchanged (and keep their dense field)."""
    a, b = tw.synth.make_pair("S", 320, 200, 8, defect=True)
    for opt in ("cfg4", "box31"):
        p = tw.OpticalFlowParameter(**OPTS[opt])
        o = tw.OpticalFlow(0, 320, 200, 1)
        want = o.calculate_batch([(a, b)], p, threshold=0.5)
        o.set_option("sparse_last", 1)
        got = o.calculate_batch([(a, b)], p, threshold=0.5)
        assert _vec(want[0]) == _vec(got[0]) and want[0]["status"] == got[0]["status"]
        o.batch_flow(0, 320, 200)
        o.close()
