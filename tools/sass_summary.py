#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libmofo_sm100.so (cuobjdump -sass): UTCHMMA (tcgen05.mma,
.2CTA = cta_group::2), UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), legacy
HMMA (mma.sync - must be 0), MUFU.EX2.   usage: tools/sass_summary.py [lib] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mofo_b200", "libmofo_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA", "MUFU.EX2", "SYNCS", "REDG"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    c = per[cur]
    c["instr"] += 1
    if op.startswith("UTCHMMA"):
        c["UTCHMMA"] += 1
        if ".2CTA" in op:
            c["UTCHMMA.2CTA"] += 1
    elif op.startswith("UTMALDG"): c["UTMALDG"] += 1
    elif op.startswith("UTMASTG"): c["UTMASTG"] += 1
    elif op.startswith("LDTM"): c["LDTM"] += 1
    elif op.startswith("STTM"): c["STTM"] += 1
    elif op.startswith("UTCBAR"): c["UTCBAR"] += 1
    elif op.startswith("HMMA"): c["HMMA"] += 1
    elif op.startswith("MUFU.EX2"): c["MUFU.EX2"] += 1
    elif op.startswith("SYNCS"): c["SYNCS"] += 1
    elif op.startswith("RED") or op.startswith("REDG"): c["REDG"] += 1
try:
    names = subprocess.run(["c++filt"] + list(per), stdout=subprocess.PIPE, text=True).stdout.splitlines()
except Exception:
    names = list(per)
print(f"# {os.path.basename(lib)}: {len(per)} kernels; columns: instr " + " ".join(KEYS))
tot = collections.Counter()
for (mangled, c), name in zip(per.items(), names):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("mofo::", "")
    print(f"{name[:64]:64s} {c['instr']:6d} " + " ".join(f"{c[k]:5d}" for k in KEYS))
    tot.update(c)
print(f"{'TOTAL':64s} {tot['instr']:6d} " + " ".join(f"{tot[k]:5d}" for k in KEYS))
