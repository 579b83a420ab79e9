#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: share of time, launches, mean.
usage: python profiles/launch_summary.py <launches.csv>"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
unit = None
for r in rows[start + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = r[ix["Metric Unit"]]
    name = r[ix["Kernel Name"]].split("(")[0][:80]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} {unit} over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / tot * 100:6.2f}%  n={v[0]:4d}  mean={v[1] / v[0]:10.1f} {unit}  {k}")
