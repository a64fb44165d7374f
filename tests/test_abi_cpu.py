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
