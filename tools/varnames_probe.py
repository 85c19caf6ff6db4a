#!/usr/bin/env python
"""C2-shape data with variable-width sample names (general path of the stats/names kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bystro_vcf_b200 import Config, Transformer, synth
N = 600000
for mode in ("fixed7", "var", "fixed10"):
    c = Config(); c.allowedFilters = {"PASS": True, ".": True}
    tr = Transformer(c)
    ch = synth.chrom_line(20130502, 2504).rstrip(b"\n").split(b"\t")
    if mode == "var": ch[9:] = [b"S%d" % (i * 7919 % 100003) for i in range(2504)]
    if mode == "fixed10": ch[9:] = [b"SAMP%06d" % i for i in range(2504)]
    tr.set_header(b"\t".join(ch))
    _, need = synth.device_lines(20130502, 2504, "chr1", 0, N, 0, 0, 0)
    d_in, _ = tr.resident_alloc(need, need // 4 + (64 << 20))
    synth.device_lines(20130502, 2504, "chr1", 0, N, d_in, need, 0)
    for _ in range(2): st, tm = tr.resident_run(need)
    best = min((tr.resident_run(need)[1] for _ in range(3)), key=lambda t: t["total_ms"])
    print(mode, {k: round(v, 3) for k, v in best.items() if k.endswith("_ms")}, "out MB", st["out_bytes"] >> 20)
    tr.close()
