#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS listing into blocks of equal execution count."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = hdr.index("Instructions Executed"); st = hdr.index("Warp Stall Sampling (All Samples)")
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
tot = 0; grp = []
for r in rows[2:]:
    c = int(r[ix]); s = int(r[st]); tot += c
    op = r[1].strip()[:44]
    if grp and grp[-1][0] == c:
        grp[-1][1] += 1; grp[-1][2] += s; grp[-1][4] = op
    else:
        grp.append([c, 1, s, op, op])
print("total warp-instructions", tot)
for g in grp:
    if g[0] * g[1] > tot * thr:
        print("%12d x %3d instr = %5.1f%%  stalls %6d | %s ... %s" % (g[0], g[1], 100 * g[0] * g[1] / tot, g[2], g[3], g[4]))
