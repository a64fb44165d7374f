#!/usr/bin/env python
"""Hot spots of the kernels in an .ncu-rep captured with --import-source on (read here, no GPU): per kernel the SASS lines
with the most warp-stall samples and their stall reasons.  usage: tools/ncu_hot.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs, cur = [], None
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == "Kernel Name":
        cur = [r[1], None, []]; secs.append(cur); continue
    if r and r[0] == "Address":
        cur[1] = r; continue
    if cur and cur[1]:
        cur[2].append(r)
for name, hdr, body in secs:
    iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in body if len(r) > iE and r[iN].isdigit()]
    tot = sum(int(r[iN]) for r in body)
    print(name[:90], "| samples", tot, "| SASS lines", len(body), "| warp-inst executed %.1fM" % (sum(int(r[iE]) for r in body) / 1e6))
    for k in sorted(range(len(body)), key=lambda k: -int(body[k][iN]))[:topn]:
        r = body[k]
        ss = ", ".join(f"{hdr[i][6:]}={r[i]}" for i in st if r[i] not in ("", "0"))
        print("  %5d %6s %9s  %-62s %s" % (k, r[iN], r[iE], r[iS].strip()[:62], ss[:90]))
    # regions between EXIT instructions (the warp roles of a specialised kernel end in EXIT)
    import collections
    a = 0
    for k in range(len(body) + 1):
        if k == len(body) or body[k][iS].strip().split()[-1].startswith("EXIT") or " EXIT" in body[k][iS]:
            n = sum(int(body[j][iN]) for j in range(a, min(k + 1, len(body)))); e = sum(int(body[j][iE]) for j in range(a, min(k + 1, len(body))))
            c = collections.Counter()
            for j in range(a, min(k + 1, len(body))):
                for i in st:
                    if body[j][i] not in ("", "0"): c[hdr[i][6:]] += int(body[j][i])
            if n: print("  region %d-%d: samples %d (%.1f%%), executed %.1fM: %s" % (a, k, n, 100 * n / max(tot, 1), e / 1e6, dict(c.most_common(7))))
            a = k + 1
