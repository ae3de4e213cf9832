#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_attention.py tests/test_gpu_reference_parity.py tests/test_gpu_model.py
grep -h "sum_i dQ_i" gpurun_out/test_gpu_attention.log
for small in 1 0; do
  MOFO_ATTN_SMALL=$small PB=32 PS=160 PH=12 timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/attn160_small${small}.csv python tools/prof_attn.py > gpurun_out/attn160_small${small}.log 2>&1
  python - <<PY
import csv
lines=[l for l in open("gpurun_out/attn160_small${small}.csv") if not l.startswith("==")]
rows=list(csv.DictReader(lines))
import collections
agg=collections.OrderedDict()
for r in rows:
    k=(r["ID"], r["Kernel Name"][:40])
    agg.setdefault(k,{})[r["Metric Name"]]=r["Metric Value"]
for k,v in agg.items():
    print(k[1], {m.split(".")[0][-28:]:x for m,x in v.items()})
PY
done
