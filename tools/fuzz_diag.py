#!/usr/bin/env python
"""Replay one fuzz seed of tests/test_gpu_parity.py and print the first differing rows (GPU vs oracle)."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T

seed = int(sys.argv[1])
rng = random.Random(1000 + seed)
n_samples = rng.choice([0, 1, 3, 31, 127, 128, 129, 300, 700, 1500])
name_w = rng.choice([7, 7, 0, 4])
vcf = T._fuzz_vcf(rng, n_samples, rng.randrange(20, 120), name_w)
kw = rng.choice([{}, {"keep_info": True, "keep_id": True}, {"keep_pos": True}, {"allow": None}, {"exclude": ["q10"]}])
print("seed", seed, "n_samples", n_samples, "name_w", name_w, "kw", kw, "bytes", len(vcf))
g = T.gpu_rows(vcf, **kw).split(b"\n"); o = T.oracle_rows(vcf, **kw).split(b"\n")
print("rows", len(g), len(o))
lines = vcf.split(b"\n")
shown = 0
for i, (a, b) in enumerate(zip(g, o)):
    if a != b:
        fa, fb = a.split(b"\t"), b.split(b"\t")
        print("row", i, "differs")
        for c, (x, y) in enumerate(zip(fa, fb)):
            if x != y: print("  col", c, "gpu", x[:200], "| oracle", y[:200])
        # find the input line: match chrom(without chr)/pos
        key = fb[1]
        for ln, L in enumerate(lines):
            f = L.split(b"\t")
            if len(f) > 8 and f[1] == key:
                gts = f[9:]
                import collections
                print("  input line", ln, "prefix", b"\t".join(f[:9])[:160], "nfields", len(f), "prefix_len", len(b"\t".join(f[:9])) + 1)
                print("  gt histogram", collections.Counter(gts).most_common(12))
                off = sum(len(x) + 1 for x in lines[:ln])
                print("  line offset in file", off, "len", len(L))
        shown += 1
        if shown >= 3: break
