#!/usr/bin/env python
"""Does running two independent pipelines concurrently (scan of one overlapping the latency-bound emit kernels of
the other) beat running them back to back?  Two contexts, two host threads, one GPU."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bystro_vcf_b200 import Config, Transformer, synth

N = int(os.environ.get("LINES", "600000"))
SUB = int(os.environ.get("SUB", str(1 << 30)))
def mk(first):
    c = Config(); c.allowedFilters = {"PASS": True, ".": True}
    tr = Transformer(c, resident_subchunk_bytes=SUB)
    tr.set_header(synth.chrom_line(20130502, 2504))
    _, need = synth.device_lines(20130502, 2504, "chr1", first, N, 0, 0, 0)
    d_in, _ = tr.resident_alloc(need, need // 8 + (64 << 20))
    synth.device_lines(20130502, 2504, "chr1", first, N, d_in, need, 0)
    return tr, need
a, na = mk(0); b, nb = mk(N)
for tr, n in ((a, na), (b, nb)):
    for _ in range(3): tr.resident_run(n)
t0 = time.perf_counter()
for _ in range(5):
    a.resident_run(na); b.resident_run(nb)
seq = (time.perf_counter() - t0) / 5
def loop(tr, n, k):
    for _ in range(k): tr.resident_run(n)
t0 = time.perf_counter()
ts = [threading.Thread(target=loop, args=(a, na, 5)), threading.Thread(target=loop, args=(b, nb, 5))]
[t.start() for t in ts]; [t.join() for t in ts]
par = (time.perf_counter() - t0) / 5
print("lines/ctx", N, "sub", SUB, "sequential ms", seq * 1e3, "concurrent ms", par * 1e3, "gain", seq / par)
