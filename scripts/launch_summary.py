#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel. Usage: launch_summary.py file.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    name = r[ki].split("(")[0].replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v for _, v in agg.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{k}` | {n} | {v / 1e3:.3f} | {v / tot:.3f} |")
