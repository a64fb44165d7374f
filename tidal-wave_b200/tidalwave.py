"""JS-level API of the reference: ``create(targetDir, options)`` -> emitter of 'data' / 'error' / 'finish'.

Mirrors /root/reference/index.js:14-73 (directory pairing + lifecycle) and the option parsing of
Broker::createInstance (/root/reference/src/broker.cpp:101-123, typed defaults :106-117, wrong-typed keys silently fall back
to the default :190-209) on top of the dispatcher (tw_pool_*).  Files are decoded with ``imread_gray`` (PNG / JPEG / PGM,
bit-exact to cv::imread IMREAD_GRAYSCALE); a file that cannot be decoded becomes the reference's "Can't open <path>" error.
"""
from __future__ import annotations

import glob
import os

from .api import OpticalFlowParameter, Pool, imread_gray


def _number(opts, key, default):
    v = opts.get(key, default)
    return float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else default  # getNumberOrDefault


def _int32(opts, key, default):
    v = opts.get(key, default)
    return int(v) if isinstance(v, int) and not isinstance(v, bool) else default  # getInt32OrDefault


class TidalWave:
    """The addon's ``TidalWave`` class (src/broker.cpp:29-42): ``calc(expected, target)``, ``dispose()``, events."""

    def __init__(self, options=None, devices=None, batch=8):
        o = options or {}
        self.threshold = _number(o, "threshold", 5.0)
        self.span = _int32(o, "span", 10)
        self.numThreads = _int32(o, "numThreads", 4)
        self.param = OpticalFlowParameter(
            pyrScale=_number(o, "pyrScale", 0.5), pyrLevels=_int32(o, "pyrLevels", 3), winSize=_int32(o, "winSize", 30),
            pyrIterations=_int32(o, "pyrIterations", 3), polyN=_int32(o, "polyN", 7), polySigma=_number(o, "polySigma", 1.5),
            flags=_int32(o, "flags", 256))
        self._handlers = {"data": [], "error": [], "finish": []}
        self._pending = []   # (request id | None, expect path, target path, immediate error | None)
        self._report = {"request": 0, "data": 0, "error": 0}
        self._devices = devices
        self._batch = batch
        self._pool = None
        self._disposed = False

    # EventEmitter
    def on(self, event, fn):
        self._handlers[event].append(fn)
        return self

    def _emit(self, event, payload):
        for fn in self._handlers[event]:
            fn(payload)

    def _ensure_pool(self, w, h):
        if self._pool is None:
            import ctypes as C
            from .api import load
            ndev = max(1, load().tw_device_count())
            devices = self._devices or [i % ndev for i in range(max(1, self.numThreads))]  # consumer i <-> GPU i % count
            # screenshot directories mix page sizes: no size bound, and every vector of every pair comes back whatever its size
            self._pool = Pool(devices, self.param, self.threshold, self.span, max_w=0, max_h=0, batch=self._batch, vector_cap=0)
            load().tw_pool_set_decoders(self._pool.pool, max(1, self.numThreads))
        return self._pool

    def calc(self, expected: str, target: str):
        """Broker::requestCalc -> Manager::request (src/broker.cpp:125-150, src/manager.cpp:68-78)."""
        if self._disposed:
            return
        self._report["request"] += 1
        # OpticalFlow::calculate's checks, in its order (src/opticalflow.cpp:26-49)
        if not expected:
            self._pending.append((None, expected, target, "ExpectImagePath is empty.")); return
        if not target:
            self._pending.append((None, expected, target, "TargetImagePath is empty.")); return
        # Manager::request with the two paths: the imreads run on the dispatcher's C++ decoder threads (numThreads of them, as the
        # reference decodes inside its numThreads consumers, src/consumer.cpp:54), the pair joins the compute queue when both are in
        pool = self._ensure_pool(0, 0)
        self._pending.append((pool.request_files(expected, target), expected, target, None))

    def flush(self):
        """Delivers every outstanding answer as 'data' / 'error' events (the uv_async hop of src/manager.cpp:102-125)."""
        pending, self._pending = self._pending, []
        queued = pending  # (request id | None, expected, target, error message | None), in request order
        for rid, expected, target, err in queued:
            if err is None:
                r = self._pool.wait(rid)
                if r is None:
                    continue  # dropped by dispose()
                if r["status"] == "ERROR":
                    err = r["reason"]
            if err is not None:
                self._report["error"] += 1
                self._emit("error", {"status": "ERROR", "reason": err})  # src/broker.cpp:57-70
            else:
                self._report["data"] += 1
                if r.pop("n_vectors", len(r["vector"])) != len(r["vector"]):
                    raise RuntimeError("dispatcher returned a truncated vector list")  # cannot happen with vector_cap = 0
                r["expect_image"], r["target_image"] = expected, target
                self._emit("data", r)                                    # src/broker.cpp:44-55,161-188

    def dispose(self):
        """Broker::requestDispose -> Manager::stop -> 'finish' with the Report (src/broker.cpp:72-86,152-158)."""
        if self._disposed:
            return
        self.flush()
        self._disposed = True
        if self._pool is not None:
            self._pool.stop()
            self._pool.close()
            self._pool = None
        self._emit("finish", dict(self._report))


def create(targetDir, options=None, devices=None, batch=8) -> TidalWave:
    """index.js:14-31.  ``options`` must carry ``expectDir`` (string) or ``getExpectedPath`` (callable)."""
    options = options or {}
    t = TidalWave(options, devices=devices, batch=batch)
    if isinstance(options.get("expectDir"), str):
        def getExpectedPath(shortPath):
            return os.path.abspath(os.path.join(options["expectDir"], shortPath))
    elif callable(options.get("getExpectedPath")):
        getExpectedPath = options["getExpectedPath"]
    else:
        raise ValueError('An option must have "expectDir" or "getExpectedPath" property.')
    t._calc_all = lambda: _calc_all(t, targetDir, getExpectedPath)
    return t


def _calc_all(t: TidalWave, targetDir, getExpectedPath):
    """index.js:33-73: glob targetDir/**/*.*, map each file to its expected file, calc, dispose when all answered."""
    base = os.path.abspath(targetDir)
    for target in sorted(glob.glob(os.path.join(base, "**", "*.*"), recursive=True)):
        if not os.path.isfile(target):
            continue
        expected = getExpectedPath(os.path.relpath(target, base))
        if not expected:
            continue  # index.js:50-51
        t.calc(expected, target)  # FS.exists' answer is ignored by the reference too (index.js:57-58)
    t.dispose()
    return t


def run(t: TidalWave) -> dict:
    """Drives the pairing + lifecycle of a ``create()``d instance to completion; returns the final Report."""
    t._calc_all()
    return dict(t._report)
