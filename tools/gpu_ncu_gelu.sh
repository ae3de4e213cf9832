#!/bin/bash
# ncu --set full captures (with source) of the decoder MLP GEMMs: fc1 + GELU, fc2 dgrad + GELU backward, fc2 + residual
mkdir -p gpurun_out
for c in dec_fc1_gelu dec_fc2_dgrad_gelubwd dec_fc2_resid; do
ncu --set full --clock-control none --import-source on -k regex:gemm_tn2_kernel -s 2 -c 1 -o gpurun_out/r02_$c -f \
  env ONLY=$c N=2 python tools/prof_gemm.py > gpurun_out/ncu_$c.log 2>&1
tail -1 gpurun_out/ncu_$c.log
done
