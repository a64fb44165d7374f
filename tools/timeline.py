#!/usr/bin/env python
"""Per-CTA phase timeline of the level-0 window kernel (development build: `make -C tidal-wave_b200/csrc dev`).
usage (GPU box): TW_LIB=tidal-wave_b200/libtidalwave_b200_dev.so python tools/timeline.py [batch] > gpurun_out/timeline.json
Each CTA records %globaltimer at: start of phase V, end of V, end of H, end of U, plus its SM id.  Prints a summary
(phase durations, per-SM overlap of the two resident CTAs) and dumps the raw records."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tidalwave_b200 as tw

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # development experiments (g_dev_flags in tw_kernels.cu)
pairs = [tw.synth.make_pair("S" if i % 2 else "T", 1920, 1080, 100 + i, defect=(i % 8 == 0)) for i in range(min(batch, 4))]
pairs = (pairs * batch)[:batch]
of = tw.OpticalFlow(0, 1920, 1080, batch)
of.set_option("graph", 0)
assert tw.load().tw_debug_set_flags(flags) == 0
for _ in range(3):
    of.calculate_batch(pairs)
lib = tw.load()
n = 20 * 34 * batch
buf = (C.c_ulonglong * (n * 24))()
rc = lib.tw_debug_timeline(buf, n * 24)
assert rc == 0, rc
t = np.frombuffer(buf, dtype=np.uint64).reshape(n, 24).astype(np.int64)
t0 = t[:, 0].min()
# slots: 0 start of V, 1 walker done (thread 0), 2 V done (barrier), 3 H done (barrier), 4/5/6 U chunk 0/1/2 done (thread 0), 7 end
rec = np.concatenate([t[:, 23:24], t[:, 0:8] - t0], 1)
sub = t[:, 8:20] - t0  # per chunk k: [8+4k] after the vote, [9+4k] all loads arrived (0 where the per-pixel path ran)
def st(d):
    return [float(np.median(d)), float(d.mean()), float(np.percentile(d, 90))]
out = {"flags": flags, "batch": batch, "ctas": int(n), "kernel_ns": int(rec[:, 8].max()),
       "V_walker_ns": st(rec[:, 2] - rec[:, 1]), "V_h2item_ns": st(rec[:, 3] - rec[:, 2]), "V_ns": st(rec[:, 3] - rec[:, 1]),
       "H_ns": st(rec[:, 4] - rec[:, 3]), "U_chunk0_ns": st(rec[:, 5] - rec[:, 4]), "U_chunk1_ns": st(rec[:, 6] - rec[:, 5]),
       "U_chunk2_ns": st(rec[:, 7] - rec[:, 6]), "U_ns": st(rec[:, 8] - rec[:, 4]), "life_ns": st(rec[:, 8] - rec[:, 1])}
ok = (t[:, 8] > 0) & (t[:, 9] > 0) & (t[:, 12] > 0) & (t[:, 13] > 0) & (t[:, 16] > 0) & (t[:, 17] > 0)
if ok.any():
    r_, s_ = rec[ok], sub[ok]
    starts = [r_[:, 4], r_[:, 5], r_[:, 6]]; ends = [r_[:, 5], r_[:, 6], r_[:, 7]]
    out["U_chunk_breakdown_ns (flow->vote, vote->loads arrived, arrived->stores issued)"] = [
        [st(s_[:, 4 * k] - starts[k])[0], st(s_[:, 4 * k + 1] - s_[:, 4 * k])[0], st(ends[k] - s_[:, 4 * k + 1])[0]] for k in range(3)]
    out["fast_path_ctas"] = int(ok.sum())
rec = rec[:, [0, 1, 3, 4, 8]]
# per SM: fraction of time with 0 / 1 / 2 CTAs in phase U, and in V|H
grid = np.arange(0, rec[:, 4].max(), 200)
occ = {"U2": 0, "U1": 0, "U0": 0}
tot = 0
for sm in np.unique(rec[:, 0]):
    r = rec[rec[:, 0] == sm]
    inU = ((grid[None, :] >= r[:, 3:4]) & (grid[None, :] < r[:, 4:5])).sum(0)
    live = ((grid[None, :] >= r[:, 1:2]) & (grid[None, :] < r[:, 4:5])).sum(0)
    m = live == 2
    occ["U2"] += int((inU[m] == 2).sum()); occ["U1"] += int((inU[m] == 1).sum()); occ["U0"] += int((inU[m] == 0).sum()); tot += int(m.sum())
out["both_resident_fraction_in_U"] = {k: v / max(tot, 1) for k, v in occ.items()}
sm0 = rec[rec[:, 0] == rec[0, 0]]
out["sm_example"] = sm0[np.argsort(sm0[:, 1])][:12].tolist()
print(json.dumps(out))
np.save(os.path.join(ROOT, "gpurun_out", "timeline_raw.npy"), rec)
