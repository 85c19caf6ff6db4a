#!/usr/bin/env python
"""Where does an end-to-end step of the slot pipeline (bvcf_submit / bvcf_collect, pinned host chunks) spend its time?
SHAPE=sites|chr1 LINES=.. CHUNK_MB=.. : per-chunk wall times with one chunk in flight, then with n_slots in flight."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bystro_vcf_b200 import Config, Transformer, synth, _lib

shape = os.environ.get("SHAPE", "sites")
ns = 0 if shape == "sites" else 2504
seed = 50 if shape == "sites" else 20130502
N = int(os.environ.get("LINES", "12000000" if shape == "sites" else "200000"))
chunk = int(os.environ.get("CHUNK_MB", "128")) << 20
L = _lib.lib()
c = Config(); c.allowedFilters = {"PASS": True, ".": True}
tr = Transformer(c, max_chunk_bytes=2 * chunk + (64 << 20))
tr.set_header(synth.chrom_line(seed, ns))
_, need = synth.device_lines(seed, ns, shape, 0, N, 0, 0, 0)
d_in, _ = tr.resident_alloc(need, need // 2 + (64 << 20))
synth.device_lines(seed, ns, shape, 0, N, d_in, need, 0)
hp = C.c_void_p()
_lib.check(L.bvcf_host_alloc(C.byref(hp), need), None, "host_alloc")
tr.resident_peek(0, need, hp.value)
hv = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(need,))
cuts = [0]
while cuts[-1] < need:
    e = min(need, cuts[-1] + chunk)
    if e < need:
        e = cuts[-1] + int(np.flatnonzero(hv[cuts[-1]:e] == 10)[-1]) + 1
    cuts.append(e)
print("bytes", need, "chunks", len(cuts) - 1, flush=True)
tsv, n, st = C.c_void_p(), C.c_size_t(), _lib.CChunkStats()

def run(depth, verbose):
    sub = col = 0
    nch = len(cuts) - 1
    t0 = time.perf_counter()
    ts = {}
    while col < nch:
        while sub < nch and sub - col < depth:
            ts[sub] = time.perf_counter()
            _lib.check(L.bvcf_submit(tr._ctx, sub, hp.value + cuts[sub], cuts[sub + 1] - cuts[sub]), tr._ctx, "submit")
            sub += 1
        tc = time.perf_counter()
        _lib.check(L.bvcf_collect(tr._ctx, col, C.byref(tsv), C.byref(n), None, None, None, C.byref(st)), tr._ctx, "collect")
        t1 = time.perf_counter()
        if verbose and col < 12:
            print("  chunk %d: submit->collected %.2f ms, in collect %.2f ms, out %d MB, retries %d" %
                  (col, (t1 - ts[col]) * 1e3, (t1 - tc) * 1e3, n.value >> 20, st.retries))
        _lib.check(L.bvcf_release(tr._ctx, col), tr._ctx, "release")
        col += 1
    return (time.perf_counter() - t0) * 1e3

for depth in (1, tr.n_slots):
    run(depth, False)
    run(depth, False)
    ms = run(depth, True)
    print("depth %d: %.1f ms per pass = %.1f GB/s of input" % (depth, ms, need / ms / 1e6), flush=True)
