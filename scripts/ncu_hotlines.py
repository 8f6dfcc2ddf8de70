#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python scripts/ncu_hotlines.py src.csv [N]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if len(r) == 2:
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or not r[0].isdigit():
        continue
    i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    a = agg[(cur_file, int(r[0]), r[1].strip()[:100])]
    num = lambda x: int(x) if x.strip().isdigit() else 0
    a[0] += num(r[i_i]); a[1] += num(r[i_t]); a[2] += num(r[i_s])
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values()); tott = sum(v[1] for v in agg.values())
print(f"total warp-instructions {tot}, thread-instructions {tott} ({tott / max(tot, 1):.2f} threads/inst), samples {tots}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:18s}:{k[1]:4d} inst {100 * v[0] / tot:5.1f}% samp {100 * v[2] / max(tots, 1):5.1f}% thr/inst {v[1] / max(v[0], 1):5.1f} | {k[2]}")
