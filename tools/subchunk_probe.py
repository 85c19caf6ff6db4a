#!/usr/bin/env python
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bystro_vcf_b200 import Config, Transformer, synth
N = int(os.environ.get("LINES", "3000000"))
for sub_gib in [int(x) for x in os.environ.get("SUB_GIB", "16,20,24").split(",")]:
    c = Config(); c.allowedFilters = {"PASS": True, ".": True}
    tr = Transformer(c, resident_subchunk_bytes=sub_gib << 30)
    tr.set_header(synth.chrom_line(20130502, 2504))
    _, need = synth.device_lines(20130502, 2504, "chr1", 0, N, 0, 0, 0)
    d_in, _ = tr.resident_alloc(need, need // 8 + (64 << 20))
    synth.device_lines(20130502, 2504, "chr1", 0, N, d_in, need, 0)
    for _ in range(2): tr.resident_run(need)
    best = min(tr.resident_run(need)[1]["total_ms"] for _ in range(4))
    print("subchunk GiB", sub_gib, "total ms", round(best, 3), "Mvar/s", round(N / best / 1e3, 1))
    tr.close()
