#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output: warp instructions per CUDA source line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None; hdr = None; agg = collections.Counter(); stl = collections.Counter(); src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ix = r.index("Instructions Executed"); sx = r.index("Warp Stall Sampling (All Samples)"); continue
    if r[0] in ("Function Name",) or hdr is None: continue
    if r[0] != "" and r[2] == "-":
        try:
            agg[(cur, int(r[0]))] += int(r[ix]); stl[(cur, int(r[0]))] += int(r[sx]); src[(cur, int(r[0]))] = r[1].strip()
        except ValueError:
            pass
tot = sum(agg.values()); ts = sum(stl.values())
print("total warp instr", tot, "stall samples", ts)
for k, v in agg.most_common(top):
    print("%5.1f%% %5.1f%%st %-18s %4d | %s" % (100 * v / tot, 100 * stl[k] / max(ts, 1), k[0], k[1], src[k][:100]))
