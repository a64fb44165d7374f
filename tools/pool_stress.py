#!/usr/bin/env python
"""Dispatcher stress (B200): image- and path-based requests of three sizes plus error cases, submitted from four threads into one pool
(two consumers, pipelined batches, three decoder threads); every answer is checked against a single synchronous context; then pools
that are stopped with work queued / in flight must neither hang nor answer wrongly.  usage: tools/pool_stress.py [n=400]"""
import os, sys, tempfile, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidalwave_b200 as tw

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
sizes = [(320, 200), (200, 320), (640, 360)]
cases = []
d = tempfile.mkdtemp(prefix="tw_stress_")
o = tw.OpticalFlow(0, 0, 0, 1)
for k, (w, h) in enumerate(sizes):
    for j in range(3):
        a, b = tw.synth.make_pair("S" if j else "T", w, h, 10 * k + j, defect=(j == 2))
        pa, pb = os.path.join(d, f"a{k}{j}.pgm"), os.path.join(d, f"b{k}{j}.pgm")
        for p, im in ((pa, a), (pb, b)):
            with open(p, "wb") as f:
                f.write(b"P5\n%d %d\n255\n" % (im.shape[1], im.shape[0]) + im.tobytes())
        want = o.calculate(a, b, threshold=0.5)
        cases.append((a, b, pa, pb, (want["status"], [(v["x"], v["y"], v["dx"], v["dy"]) for v in want["vector"]])))
o.close()
pool = tw.Pool([0, 0], batch=4, threshold=0.5)
tw.load().tw_pool_set_decoders(pool.pool, 3)
ids, lock = [None] * n, threading.Lock()

def submit(t):
    rng = np.random.default_rng(t)
    for i in range(t, n, 4):
        c = cases[int(rng.integers(len(cases)))]
        kind = int(rng.integers(10))
        with lock:  # the Python wrapper's bookkeeping is not thread-safe; the C calls are
            if kind == 0:
                ids[i] = (pool.request_files(c[2], os.path.join(d, "missing.pgm")), ("ERROR", None))
            elif kind == 1:
                ids[i] = (pool.request(c[0], np.zeros((c[0].shape[0] + 9, c[0].shape[1]), np.uint8)), ("ERROR", None))
            elif kind < 6:
                ids[i] = (pool.request_files(c[2], c[3]), c[4])
            else:
                ids[i] = (pool.request(c[0], c[1]), c[4])
ths = [threading.Thread(target=submit, args=(t,)) for t in range(4)]
t0 = time.time()
[t.start() for t in ths]; [t.join() for t in ths]
bad = 0
for rid, (st, vec) in ids:
    r = pool.wait(rid)
    got = (r["status"], [(v["x"], v["y"], v["dx"], v["dy"]) for v in r.get("vector", [])])
    if got[0] != st or (vec is not None and got[1] != vec):
        bad += 1
rep = pool.report()
pool.stop(); pool.close()
print("stress: %d requests in %.2f s, mismatches %d, report %s" % (n, time.time() - t0, bad, rep))
assert bad == 0 and rep["request"] == n
for trial in range(5):  # stop with work queued / in flight
    pool = tw.Pool([0], batch=4, threshold=0.5)
    c = cases[trial % len(cases)]
    rids = [pool.request_files(c[2], c[3]) if i % 2 else pool.request(c[0], c[1]) for i in range(120)]
    time.sleep(0.002 * trial)
    pool.stop()
    res = [pool.wait(r) for r in rids]
    done = [r for r in res if r is not None]
    assert all(r["status"] == c[4][0] for r in done), "wrong answer after stop"
    pool.close()
    print("stop trial %d: %d answered, %d dropped" % (trial, len(done), len(res) - len(done)), flush=True)
print("ok")
