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
print(of.calculate(a, np.ascontiguousarray(b[:147, :196]))["status"])
print([r["status"] for r in of.calculate_batch([(a, b), (a, a)])])
of.close()
pool = tw.Pool([0, 0], batch=2, max_w=200, max_h=150)
ids = [pool.request(a, b) for _ in range(5)]
print([pool.wait(i)["status"] for i in ids], pool.report())
pool.stop(); pool.close()
print("done")
