#!/usr/bin/env python
"""BASELINE.json configs[4]: 10,000 synthetic 1920x1080 pairs through the in-process per-GPU work queue (tw_pool_*: one
process, consumer threads bound to every visible GPU, like the reference's Manager + Consumers), results checked against a
single-context run.  usage: tools/config5.py [n_pairs] [consumers_per_gpu] [batch]"""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tidalwave_b200 as tw

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
CPG = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
W, H, POOL = 1920, 1080, 64
lib = tw.load()
ngpu = lib.tw_device_count()
pairs = tw.synth.pool_pairs(POOL, W, H, seed0=100)
pinned = []
for a, b in pairs:
    pa = lib.tw_host_alloc(W * H); pb = lib.tw_host_alloc(W * H)
    C.memmove(pa, a.ctypes.data, W * H); C.memmove(pb, b.ctypes.data, W * H)
    pinned.append((pa, pb))
# reference answers from one context
of = tw.OpticalFlow(0, W, H, 8)
want = []
for i in range(0, POOL, 8):
    want += [(r["status"], [(v["x"], v["y"], v["dx"], v["dy"]) for v in r["vector"]]) for r in of.calculate_batch(pairs[i:i + 8])]
of.close()
devices = [g for g in range(ngpu) for _ in range(CPG)]
p = tw.OpticalFlowParameter().c()
err = C.create_string_buffer(256)
pool = lib.tw_pool_create((C.c_int * len(devices))(*devices), len(devices), W, H, B, C.byref(p), 5.0, 10, 4096, err, 256)
assert pool, err.value
vec = (tw.tw_vector * 4096)(); res = tw.tw_result()
def run(n):
    ids = [lib.tw_pool_submit(pool, pinned[i % POOL][0], W, H, pinned[i % POOL][1], W, H) for i in range(n)]
    bad = 0
    for i, rid in enumerate(ids):
        rc = lib.tw_pool_wait(pool, rid, vec, 4096, C.byref(res))
        st, wv = want[i % POOL]
        got = [(vec[k].x, vec[k].y, vec[k].dx, vec[k].dy) for k in range(min(res.n_vectors, 4096))]
        if rc != 0 or tw.api.STATUS_NAMES[res.status] != st or got != wv:
            bad += 1
    return bad
run(len(devices) * B * 2)
t0 = time.perf_counter(); bad = run(N); dt = time.perf_counter() - t0
a = C.c_int(); b = C.c_int(); c = C.c_int()
lib.tw_pool_report(pool, C.byref(a), C.byref(b), C.byref(c))
lib.tw_pool_destroy(pool)
print(json.dumps({"config": "configs[4]", "pairs": N, "gpus": ngpu, "consumers_per_gpu": CPG, "batch": B, "pairs_per_s": N / dt, "seconds": dt,
                  "mismatches_vs_single_context": bad, "suspicious_in_pool": sum(1 for s, _ in want if s == "SUSPICIOUS"),
                  "report": {"request": a.value, "data": b.value, "error": c.value}}))
