#!/usr/bin/env python
"""Compact per-launch summary of an `ncu --set full` report (run where ncu is installed; no GPU needed):
duration, DRAM bytes read / written and the achieved DRAM GB/s against the measured copy peak (MEASURED_PEAKS.json),
tensor-pipe / issue / XU utilisation, occupancy, registers.   usage: tools/ncu_summary.py report.ncu-rep > profiles/x.txt"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
peak = 6516.7
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}


def val(r, k):
    if k not in ix or r[ix[k]] in ("", "n/a"):
        return None
    return float(r[ix[k]].replace(",", "")) * SCALE.get(units[ix[k]], 1.0)


print(f"# {os.path.basename(rep)}: ncu --set full --clock-control none; DRAM peak for the fraction = measured copy bandwidth {peak:.1f} GB/s")
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    dur, rd, wr = val(r, "gpu__time_duration.sum"), val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / dur / 1e9 if dur else 0.0
    g = lambda k: r[ix[k]] if k in ix else "n/a"
    print(f"{name[:58]:58s} grid {g('Grid Size'):>16s} block {g('Block Size'):>14s} regs {g('launch__registers_per_thread'):>4s} | {dur * 1e6:8.1f} us | "
          f"dram read {rd / 1e6:9.2f} MB write {wr / 1e6:9.2f} MB -> {gbs:7.1f} GB/s = {gbs / peak:5.2f} of peak | "
          f"tensor {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):>6s}% issue {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):>6s}% "
          f"xu {g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):>6s}% warps {g('sm__warps_active.avg.pct_of_peak_sustained_active'):>6s}% "
          f"smem-conflicts {g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')}")
