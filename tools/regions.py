#!/usr/bin/env python
"""Bucket the warp instructions / stall samples of an `ncu --page source --csv --print-source cuda,sass` dump by
source regions: python tools/regions.py dump.csv file:lo-hi=name ... (SASS is walked in address order; lines outside
every region inherit the previous instruction's region)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
regs = []
for a in sys.argv[2:]:
    spec, name = a.split("=")
    f, rng = spec.split(":")
    lo, hi = rng.split("-")
    regs.append((f, int(lo), int(hi), name))
hdr = None; cur = None; curline = None; amap = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; ix = r.index("Instructions Executed"); sx = r.index("Warp Stall Sampling (All Samples)"); continue
    if hdr is None or len(r) < 4: continue
    if r[0] != "" and r[2] == "-":
        try: curline = (cur, int(r[0]))
        except ValueError: curline = None
    elif r[0] == "" and r[2].startswith("0x"):
        amap[int(r[2], 16)] = (curline, int(r[ix]), int(r[sx]))
def region(fl):
    for f, lo, hi, name in regs:
        if fl[0] == f and lo <= fl[1] <= hi: return name
    return None
reg = collections.Counter(); st = collections.Counter(); cur = "prologue"
for a in sorted(amap):
    fl, c, s = amap[a]
    if fl:
        r = region(fl)
        if r: cur = r
    reg[cur] += c; st[cur] += s
tot = sum(reg.values()); ts = sum(st.values())
print("total warp instr %d" % tot)
for k, v in reg.most_common(): print("%-24s instr %5.1f%%  stalls %5.1f%%  %12d" % (k, 100 * v / tot, 100 * st[k] / max(ts, 1), v))
