#!/usr/bin/env python
"""Device-resident timing of the reference's own fixture (BASELINE configs[0]: 1000G chr1, 19,747 lines x 2,504 samples),
tiled N times so that the kernels see more than one wave."""
import gzip, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bystro_vcf_b200 import Config, Transformer, parse_preamble

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 8
vcf = gzip.open(os.path.join(ROOT, "tests", "golden", "chr1_20klines.vcf.gz")).read()
w, chrom, off = parse_preamble(vcf)
body = vcf[off:] * tiles
c = Config(); c.allowedFilters = {"PASS": True, ".": True}
with Transformer(c, eol_width=w) as tr:
    tr.set_header(chrom)
    tr.resident_alloc(len(body), len(body) // 6 + (64 << 20))
    step = 256 << 20
    for o in range(0, len(body), step):
        tr.resident_upload(o, body[o:o + step])
    for _ in range(3): stats, times = tr.resident_run(len(body))
    best = min((tr.resident_run(len(body))[1] for _ in range(5)), key=lambda t: t["total_ms"])
print(json.dumps({"config": "C1 fixture x%d" % tiles, "lines": stats["n_lines"], "in_GB": len(body) / 1e9, "out_GB": stats["out_bytes"] / 1e9,
                  "ms": {k: round(v, 3) for k, v in best.items() if k.endswith("_ms")},
                  "variants_per_s": stats["n_lines"] / best["total_ms"] * 1e3,
                  "scan_GBps": len(body) / best["scan_ms"] / 1e6, "pipeline_GBps": (len(body) + stats["out_bytes"]) / best["total_ms"] / 1e6}))
