#!/usr/bin/env python
"""Sum an ncu_hotlines listing by named source regions of one file.
    python scripts/ncu_regions.py hot.txt rtw_kernels.cu name:lo-hi name:lo-hi ...   (other files are listed by file name)"""
import re
import sys

hot, fname = sys.argv[1], sys.argv[2]
regs = []
for a in sys.argv[3:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    regs.append((n, int(lo), int(hi)))
agg = {}
for ln in open(hot):
    m = re.match(r"(\S+)\s*:\s*(\d+) inst\s+([\d.]+)% samp\s+([\d.]+)% thr/inst\s+([\d.]+)", ln)
    if not m:
        continue
    f, l, i, s, t = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))
    k = f
    if f == fname:
        k = "other:" + fname
        for n, lo, hi in regs:
            if lo <= l <= hi:
                k = n
                break
    a = agg.setdefault(k, [0.0, 0.0, 0.0])
    a[0] += i; a[1] += s; a[2] += i * t
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} inst {v[0]:5.1f}%  samples {v[1]:5.1f}%  lanes/inst {v[2] / max(v[0], 1e-9):5.1f}")
