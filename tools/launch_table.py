#!/usr/bin/env python
"""Per-launch table of one DiT block from an `ncu --metrics gpu__time_duration.sum --csv` launch list (diagnostics)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
names = [(r[ki], r[gi], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:]]
agg = collections.OrderedDict()
for k, g, t in names:
    a = agg.setdefault((k.split("::")[-1][:48], g), [0, 0.0])
    a[0] += 1
    a[1] += t
print(f"total {sum(v[1] for v in agg.values()):.1f} us over {len(names)} launches")
for (k, g), v in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
    print(f"{k:50s} {g:16s} n={v[0]:4d}  {v[1] / v[0]:8.1f} us/launch  {v[1]:9.1f} us")
idx = [i for i, n in enumerate(names) if "qkv_head_scatter" in n[0]]
if len(idx) > 10:
    s = idx[10] - 2
    print("-- one block (launch order)")
    for k, g, t in names[s:s + (idx[11] - idx[10])]:
        print(f"{k.split('::')[-1][:48]:50s} {g:16s} {t:8.1f} us")
