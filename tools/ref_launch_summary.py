#!/usr/bin/env python
"""Class-level summary of an ncu launch list of the reference's eager step (tools/ref_step.py): the LAST step's kernels
grouped into GEMM (cuBLAS / cutlass / nvjet), batched attention matmuls, softmax, LayerNorm, conv (patch embed), element-wise /
copy / cast, reductions, optimizer.   usage: tools/ref_launch_summary.py launches.csv n_steps"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = []
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    rows.append((r["Kernel Name"], v))
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
per = len(rows) // n_steps
step = rows[-per:]
def cls(n):
    l = n.lower()
    if "softmax" in l: return "softmax fwd/bwd"
    if "layer_norm" in l or "layernorm" in l: return "LayerNorm fwd/bwd"
    if "conv" in l or "cudnn" in l or "implicit" in l or "wgrad" in l or "dgrad" in l: return "conv3d patch embed (cuDNN)"
    if "gemm" in l or "nvjet" in l or "cutlass" in l or "cublas" in l or "sm90" in l or "sm100" in l or "xmma" in l: return "GEMM (cuBLAS: linear layers + attention bmm)"
    if "adam" in l or "multi_tensor" in l or "foreach" in l: return "optimizer / foreach (AdamW, unscale, norm)"
    if "gelu" in l: return "GELU fwd/bwd"
    if "reduce" in l or "norm" in l or "sum" in l or "mean" in l: return "reductions (mean/var/sum/norm)"
    if "index" in l or "gather" in l or "scatter" in l or "nonzero" in l or "cat" in l: return "indexing / gather / cat"
    return "element-wise / copy / cast"
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in step:
    agg[cls(n)][0] += 1; agg[cls(n)][1] += v
tot = sum(v for _, v in step)
print(f"# reference eager step (last of {n_steps}): {len(step)} kernel launches, {tot / 1e3:.2f} ms of kernel time")
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v / 1e3:9.3f} ms {100 * v / tot:5.1f}%  x{c:5d}  {k}")
print("# top kernels")
top = collections.defaultdict(lambda: [0, 0.0])
for n, v in step:
    k = re.sub(r"\(.*", "", n)[:110]
    top[k][0] += 1; top[k][1] += v
for k, (c, v) in sorted(top.items(), key=lambda x: -x[1][1])[:25]:
    print(f"{v / 1e3:9.3f} ms  x{c:5d}  {k}")
