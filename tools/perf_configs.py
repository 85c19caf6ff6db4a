#!/usr/bin/env python
"""Device-resident timing of the non-headline BASELINE.json configs (C3 sites-only, C4 biobank slice, C5)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bystro_vcf_b200 import Config, Transformer, synth

def run(name, seed, ns, shape, n_lines, cfg_kw=None, reps=3):
    c = Config(); c.allowedFilters = {"PASS": True, ".": True}
    for k, v in (cfg_kw or {}).items(): setattr(c, k, v)
    tr = Transformer(c)
    tr.set_header(synth.chrom_line(seed, ns))
    _, need = synth.device_lines(seed, ns, shape, 0, n_lines, 0, 0, 0)
    d_in, _ = tr.resident_alloc(need, need // 2 + (64 << 20))
    got, _ = synth.device_lines(seed, ns, shape, 0, n_lines, d_in, need, 0)
    assert got == need
    for _ in range(2): stats, times = tr.resident_run(need)
    best = None
    for _ in range(reps):
        stats, times = tr.resident_run(need)
        if best is None or times["total_ms"] < best["total_ms"]: best = times
    alg = need + stats["out_bytes"]
    print(json.dumps({"config": name, "lines": n_lines, "samples": ns, "in_GB": need / 1e9, "out_GB": stats["out_bytes"] / 1e9,
                      "rows": stats["n_rows"], "ms": {k: round(v, 3) for k, v in best.items() if k.endswith("_ms")},
                      "variants_per_s": n_lines / best["total_ms"] * 1e3, "pipeline_GBps": alg / best["total_ms"] / 1e6,
                      "retries": stats["retries"]}))
    tr.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["C3", "C4", "C5"]
    if "C3" in which: run("C3 sites-only", 50, 0, "sites", int(os.environ.get("C3_LINES", "20000000")))
    if "C4" in which: run("C4 biobank 200k samples (slice)", 200000, 200000, "biobank", int(os.environ.get("C4_LINES", "4000")))
    if "C5" in which: run("C5 chr1 filters keepInfo dosage", 20130502, 2504, "chr1_filters", 600000,
                          {"keepInfo": True, "allowedFilters": None, "excludedFilters": {"LowQual": True}})
    if "C5d" in which: run("C5 + dosage", 20130502, 2504, "chr1_filters", 300000,
                           {"keepInfo": True, "allowedFilters": None, "excludedFilters": {"LowQual": True}, "dosageMatrixOutPath": "x"})
