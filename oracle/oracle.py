"""CPU ORACLE bindings (test infrastructure, NOT product code).

Two oracles for the hot path /root/reference/src/opticalflow.cpp:83-85 (cv::calcOpticalFlowFarneback)
followed by /root/reference/src/consumer.cpp:60-77 (span sampling + threshold classification):

* ``RefOracle`` -- ctypes binding of oracle/libtwref.so, the C restatement in oracle/farneback_ref.c
  (op-order-faithful to SURVEY.md App. A).  Travels to the GPU box as a built .so.
* ``cv2_flow`` -- Python OpenCV (cv2 4.13.0) ``calcOpticalFlowFarneback`` pinned with
  ``cv2.ipp.setUseIPP(False)``, ``cv2.setNumThreads(1)``: the same algorithm as the reference's OpenCV
  2.4.9 dependency (un-vendored; .travis.yml:8), newer build.  Used to pin the C restatement and to
  generate tests/golden/ (tools/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libtwref.so")


@dataclass
class FlowParam:
    """Field-for-field OpticalFlowParameter (/root/reference/src/opticalflow.h:28-36) with the defaults of
    Broker::createInstance (/root/reference/src/broker.cpp:106-117)."""
    pyrScale: float = 0.5
    pyrLevels: int = 3
    winSize: int = 30
    pyrIterations: int = 3
    polyN: int = 7
    polySigma: float = 1.5
    flags: int = 256


class _CParam(C.Structure):
    _fields_ = [("pyrScale", C.c_double), ("pyrLevels", C.c_int), ("winSize", C.c_int),
                ("pyrIterations", C.c_int), ("polyN", C.c_int), ("polySigma", C.c_double), ("flags", C.c_int)]


class _CVector(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("dx", C.c_double), ("dy", C.c_double)]


_DUMPFN = C.CFUNCTYPE(None, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int)


def build(force: bool = False) -> str:
    """Compile oracle/libtwref.so with the committed Makefile (building the checker is not using it)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(
            os.path.join(_HERE, "farneback_ref.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libtwref.so"])
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class RefOracle:
    def __init__(self):
        build()
        self.lib = C.CDLL(_LIB)
        L = self.lib
        L.twref_schedule.restype = C.c_int
        L.twref_farneback_dump.restype = C.c_int
        L.twref_sample.restype = C.c_int

    def set_relax(self, bits: int) -> None:
        """Relaxed-arithmetic variants of the restatement (farneback_ref.c, twref_set_relax): 0 = faithful App. A;
        17 = what the product's "arithmetic" = 1 runs (fmaf window taps + mixed double/float poly-exp pass)."""
        self.lib.twref_set_relax(int(bits))

    @staticmethod
    def _cparam(p: FlowParam) -> _CParam:
        return _CParam(p.pyrScale, p.pyrLevels, p.winSize, p.pyrIterations, p.polyN, p.polySigma, p.flags)

    def schedule(self, W, H, pyrScale=0.5, levels=3):
        k = (C.c_int * 16)(); w = (C.c_int * 16)(); h = (C.c_int * 16)(); ks = (C.c_int * 16)()
        sg = (C.c_double * 16)()
        n = self.lib.twref_schedule(W, H, C.c_double(pyrScale), levels, k, w, h, ks, sg)
        return [dict(k=k[i], w=w[i], h=h[i], ksize=ks[i], sigma=sg[i]) for i in range(n)]

    def gauss_kernel(self, ksize, sigma):
        out = np.zeros(ksize, np.float32)
        self.lib.twref_gauss_kernel(ksize, C.c_double(sigma), _fp(out))
        return out

    def level_image(self, img, ksize, sigma, w, h):
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        out = np.empty((h, w), np.float32)
        self.lib.twref_level_image(img.ctypes.data_as(C.c_void_p), W, H, W, ksize, C.c_double(sigma), w, h, _fp(out))
        return out

    def resize_linear(self, src, w, h):
        src = np.ascontiguousarray(src, np.float32)
        H, W = src.shape[:2]
        cn = 1 if src.ndim == 2 else src.shape[2]
        out = np.empty((h, w) if cn == 1 else (h, w, cn), np.float32)
        self.lib.twref_resize_linear(_fp(src), W, H, cn, _fp(out), w, h)
        return out

    def resize_u8(self, img, w, h):
        """The +-5 px path of OpticalFlow::calculate (/root/reference/src/opticalflow.cpp:64-68): cv::resize, INTER_LINEAR."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W = img.shape
        out = np.empty((h, w), np.uint8)
        self.lib.twref_resize_u8(img.ctypes.data_as(C.c_void_p), W, H, W, out.ctypes.data_as(C.c_void_p), w, h, w)
        return out

    def calculate(self, expect, target, p: "FlowParam" = None, span=10, threshold=5.0):
        """OpticalFlow::calculate + Consumer::run on decoded images (/root/reference/src/opticalflow.cpp:52-73,
        src/consumer.cpp:59-88) -> (status, vectors, flow or None)."""
        p = p or FlowParam()
        if abs(expect.shape[0] - target.shape[0]) > 5 or abs(expect.shape[1] - target.shape[1]) > 5:
            return "ERROR", [], None
        if expect.shape != target.shape:
            target = self.resize_u8(target, expect.shape[1], expect.shape[0])
        flow = self.farneback(expect, target, p)
        status, vec = self.sample(flow, span, threshold)
        return status, vec, flow

    def polyexp_tables(self, n, sigma):
        g = np.zeros(65, np.float32); xg = np.zeros(65, np.float32); xxg = np.zeros(65, np.float32)
        ig = np.zeros(4, np.float64)
        self.lib.twref_polyexp_tables(n, C.c_double(sigma), _fp(g), _fp(xg), _fp(xxg),
                                      ig.ctypes.data_as(C.POINTER(C.c_double)))
        return g[:n + 1], xg[:n + 1], xxg[:n + 1], ig

    def polyexp(self, I, n, sigma):
        I = np.ascontiguousarray(I, np.float32)
        h, w = I.shape
        R = np.empty((h, w, 5), np.float32)
        self.lib.twref_polyexp(_fp(I), w, h, n, C.c_double(sigma), _fp(R))
        return R

    def update_matrices(self, R0, R1, flow):
        R0 = np.ascontiguousarray(R0, np.float32); R1 = np.ascontiguousarray(R1, np.float32)
        flow = np.ascontiguousarray(flow, np.float32)
        h, w = R0.shape[:2]
        M = np.empty((h, w, 5), np.float32)
        self.lib.twref_update_matrices(_fp(R0), _fp(R1), _fp(flow), w, h, _fp(M))
        return M

    def window_kernel(self, winSize):
        k = np.zeros(256, np.float32)
        self.lib.twref_window_kernel(winSize, _fp(k))
        return k[:winSize // 2 + 1]

    def blur_solve(self, M, winSize, gaussian=True):
        M = np.ascontiguousarray(M, np.float32)
        h, w = M.shape[:2]
        flow = np.empty((h, w, 2), np.float32)
        if gaussian:
            self.lib.twref_gauss_blur_solve(_fp(M), w, h, winSize, _fp(flow), None)
        else:
            self.lib.twref_box_blur_solve(_fp(M), w, h, winSize, _fp(flow), None)
        return flow

    def farneback(self, prev, nxt, p: FlowParam = FlowParam(), dump: dict | None = None):
        """Returns flow (H, W, 2) float32.  If ``dump`` is a dict it is filled with per-stage tensors keyed
        (stage, scale_index, iter)."""
        prev = np.ascontiguousarray(prev, np.uint8); nxt = np.ascontiguousarray(nxt, np.uint8)
        assert prev.shape == nxt.shape and prev.ndim == 2
        H, W = prev.shape
        flow = np.empty((H, W, 2), np.float32)
        cp = self._cparam(p)

        def _cb(stage, s, it, data, w, h, cn):
            arr = np.ctypeslib.as_array(data, shape=(h, w, cn) if cn > 1 else (h, w)).copy()
            dump[(stage.decode(), s, it)] = arr

        cb = _DUMPFN(_cb) if dump is not None else C.cast(None, _DUMPFN)
        rc = self.lib.twref_farneback_dump(prev.ctypes.data_as(C.c_void_p), nxt.ctypes.data_as(C.c_void_p), W, H, W,
                                           C.byref(cp), _fp(flow), cb)
        if rc != 0:
            raise ValueError("twref_farneback: bad parameter")
        return flow

    def sample(self, flow, span=10, threshold=5.0, cap=1 << 20):
        """/root/reference/src/consumer.cpp:60-77 -> (status, [(x, y, dx, dy), ...])."""
        flow = np.ascontiguousarray(flow, np.float32)
        H, W = flow.shape[:2]
        buf = (_CVector * cap)()
        n = self.lib.twref_sample(_fp(flow), W, H, span, C.c_double(threshold), buf, cap)
        vec = [(buf[i].x, buf[i].y, buf[i].dx, buf[i].dy) for i in range(min(n, cap))]
        return ("OK" if n == 0 else "SUSPICIOUS"), vec


def sample_numpy(flow, span=10, threshold=5.0):
    """NumPy restatement of /root/reference/src/consumer.cpp:60-77 (float len, double compare, row-major)."""
    fx = flow[::span, ::span, 0].astype(np.float32)
    fy = flow[::span, ::span, 1].astype(np.float32)
    ln = (fx * fx).astype(np.float32) + (fy * fy).astype(np.float32)
    mask = ln.astype(np.float64) > (threshold * threshold)
    ys, xs = np.nonzero(mask)
    vec = [(int(x) * span, int(y) * span, float(fx[y, x]), float(fy[y, x])) for y, x in zip(ys, xs)]
    return ("OK" if not vec else "SUSPICIOUS"), vec


def cv2_flow(prev, nxt, p: FlowParam = FlowParam()):
    """The pinned third-party oracle: cv2.calcOpticalFlowFarneback with IPP off."""
    import cv2
    cv2.ipp.setUseIPP(False)
    cv2.setNumThreads(1)
    return cv2.calcOpticalFlowFarneback(np.ascontiguousarray(prev), np.ascontiguousarray(nxt), None, p.pyrScale,
                                        p.pyrLevels, p.winSize, p.pyrIterations, p.polyN, p.polySigma, p.flags)
