#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the LAST step
(from the last tube_mask launch to the end).  usage: tools/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    rows.append((row["Kernel Name"], v))
idx = [i for i, (n, _) in enumerate(rows) if "tube_mask" in n]
step = rows[idx[-1]:] if idx else rows
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in step:
    k = re.sub(r"\(.*", "", n)
    agg[k][0] += 1
    agg[k][1] += v
total = sum(v for _, v in step)
print(f"launches in last step: {len(step)}   total kernel time: {total / 1e3:.3f} ms")
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v:10.1f} us {100 * v / total:5.1f}%  x{c:4d}  {k[:120]}")
