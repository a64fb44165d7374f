#!/usr/bin/env python
"""Per-phase instruction mix and stall-sample shares of one kernel from an .ncu-rep captured with --import-source on
(read here, no GPU).  Phases are split at BAR.SYNC.  usage: tools/ncu_opmix.py rep.ncu-rep <pixels per launch> [kernel regex]"""
import collections, csv, io, re, subprocess, sys

rep, npx = sys.argv[1], float(sys.argv[2])
kid = ["--kernel-id", "::regex:%s:1" % sys.argv[3]] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + kid, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name": cur = [r[1], None, []]; secs.append(cur); continue
    if r and r[0] == "Address": cur[1] = r; continue
    if cur: cur[2].append(r)
name, hdr, body = secs[0]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
seg, ops, smp, st = 0, collections.defaultdict(collections.Counter), collections.Counter(), collections.defaultdict(collections.Counter)
for r in body:
    if len(r) <= iE: continue
    s = r[iS].strip(); e = int(r[iE])
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', s)
    o = m.group(2) if m else s
    o = '.'.join(o.split('.')[:2]) if o.startswith(('LD', 'ST', 'F2F', 'I2F')) else o.split('.')[0]
    ops[seg][o] += e; smp[seg] += int(r[iSm])
    for i in stall:
        if r[i] not in ("", "0"): st[seg][hdr[i][6:]] += int(r[i])
    if 'BAR.SYNC' in s: seg += 1
print(name)
tot_e = sum(sum(c.values()) for c in ops.values()); tot_s = sum(smp.values())
print("thread-instructions per pixel: %.1f" % (tot_e * 32 / npx))
for sg in sorted(ops):
    c = ops[sg]; t = sum(c.values())
    print("phase %d: %.1f thread-inst/px (%.1f %%), %.1f %% of the warp samples" % (sg, t * 32 / npx, 100 * t / tot_e, 100 * smp[sg] / max(tot_s, 1)))
    print("   mix/px: " + ", ".join("%s %.1f" % (o, v * 32 / npx) for o, v in c.most_common(14)))
    print("   stalls: " + ", ".join("%s %.0f%%" % (k, 100 * v / max(smp[sg], 1)) for k, v in st[sg].most_common(7)))
