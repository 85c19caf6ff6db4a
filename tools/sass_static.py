#!/usr/bin/env python
"""Static SASS instructions per source line of one kernel: nvdisasm -g -c <cubin> > dis.txt; sass_static.py dis.txt <kernel-substr> [top]"""
import re, sys, collections
txt = open(sys.argv[1]).read().splitlines()
want = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cnt = collections.Counter(); cur = None; on = False; total = 0
for l in txt:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m: on = want in m.group(1); cur = None; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l): cnt[cur] += 1; total += 1
print("static instructions", total, "=", round(total * 16 / 1024, 1), "KB")
src = {}
for (f, n), c in cnt.most_common(top):
    try:
        if f not in src: src[f] = open("bystro_vcf_b200/csrc/" + f).read().splitlines()
        s = src[f][n - 1].strip()[:90]
    except Exception: s = ""
    print("%5d %5.1f%% %-16s %4d | %s" % (c, 100 * c / total, f, n, s))
