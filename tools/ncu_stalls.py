#!/usr/bin/env python
"""Per-instruction stall summary of an `ncu --set full --import-source on` report (runs where ncu is installed, no GPU):
stall-reason totals, the hottest SASS instructions with their two top reasons, and the instruction mix of the hottest loop
body (instructions sharing the most common execution count).   usage: tools/ncu_stalls.py report.ncu-rep [top_n]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print("#", rows[0][1] if len(rows[0]) > 1 else rows[0])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print(f"# total samples {tot:.0f}; stall reasons:", ", ".join(f"{s[6:]} {v / tot * 100:.1f}%" for s, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
print("# hottest instructions: addr samples executed source [top reasons]")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top_n]:
    top = sorted(stalls, key=lambda s: -f(r, s))[:2]
    print(f"{r[ix['Address']][-5:]} {f(r, '# Samples'):6.0f} {f(r, 'Instructions Executed'):9.0f}  {r[ix['Source']].strip()[:72]:72s} "
          + " ".join(f"{t[6:]}={int(f(r, t))}" for t in top))
cnt = collections.Counter(r[ix["Instructions Executed"]] for r in data)
body_count = max((c for c in cnt if c not in ("0", "")), key=lambda c: cnt[c] * 1.0)
body = [r for r in data if r[ix["Instructions Executed"]] == body_count]
mix = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", r[ix["Source"]].strip()).split()[0].split(".")[0] for r in body)
print(f"# hottest loop body: {len(body)} instructions executed {body_count} times each:", ", ".join(f"{k} {v}" for k, v in mix.most_common(14)))
