#!/usr/bin/env python
"""Small run through every kernel family (default, box, generic radius/polyN/pyrScale, resize, batch, pool) for compute-sanitizer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidalwave_b200 as tw
a, b = tw.synth.make_pair("S", 200, 150, 3, defect=True)
of = tw.OpticalFlow(0, 256, 256, 2)
for kw in (dict(), dict(flags=0, winSize=15, polyN=5, polySigma=1.1), dict(winSize=20), dict(polyN=3, polySigma=0.9), dict(pyrScale=0.8, pyrLevels=2),
           dict(winSize=15, polyN=5)):
    r = of.calculate(a, b, tw.OpticalFlowParameter(**kw))
    print(kw, r["status"], len(r["vector"]))
# round 2: the strip window kernel (both arithmetics, dense + sparse last), the fused box iteration with several radii, a deep
# pyramid (long pre-blur level path), the pipelined consumer and path-based requests
for tiles in (0, 1):
    of.set_option("window_tiles", tiles)
    for arith in (0, 1):
        of.set_option("arithmetic", arith)
        for sparse in (0, 1):
            of.set_option("sparse_last", sparse)
            print("tiles", tiles, "arith", arith, "sparse", sparse, of.calculate(a, b)["status"])
of.set_option("window_tiles", 1); of.set_option("sparse_last", 0)
for kw in (dict(flags=0, winSize=2), dict(flags=0, winSize=31), dict(flags=0, winSize=33), dict(flags=0, winSize=40), dict(pyrLevels=5, pyrIterations=2)):
    r = of.calculate(a, b, tw.OpticalFlowParameter(**kw))
    print(kw, r["status"], len(r["vector"]))
c, d = tw.synth.make_pair("T", 97, 61, 5)
print("small strip", of.calculate(c, d)["status"])
print(of.calculate(a, np.ascontiguousarray(b[:147, :196]))["status"])
print([r["status"] for r in of.calculate_batch([(a, b), (a, a)])])
of.close()
pool = tw.Pool([0, 0], batch=2, max_w=200, max_h=150)
ids = [pool.request(a, b) for _ in range(5)]
print([pool.wait(i)["status"] for i in ids], pool.report())
pool.stop(); pool.close()
import tempfile
d = tempfile.mkdtemp()
np.save(os.path.join(d, "x.npy"), a)
with open(os.path.join(d, "a.pgm"), "wb") as f: f.write(b"P5\n%d %d\n255\n" % (a.shape[1], a.shape[0]) + a.tobytes())
with open(os.path.join(d, "b.pgm"), "wb") as f: f.write(b"P5\n%d %d\n255\n" % (b.shape[1], b.shape[0]) + b.tobytes())
pool = tw.Pool([0], batch=3)
ids = [pool.request_files(os.path.join(d, "a.pgm"), os.path.join(d, "b.pgm")) for _ in range(7)] + [pool.request_files(os.path.join(d, "a.pgm"), "")]
print([pool.wait(i)["status"] for i in ids], pool.report())
pool.stop(); pool.close()
print("done")
