#!/bin/bash
# one ncu --set full capture (with source) of the decoder fc1 + GELU GEMM and of the decoder attention forward / backward
mkdir -p gpurun_out
ONLY=dec_fc1_gelu N=2 python tools/prof_gemm.py > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_tn2_kernel -s 2 -c 1 -o gpurun_out/r02_gelu_v2 -f \
  env ONLY=dec_fc1_gelu N=2 python tools/prof_gemm.py > gpurun_out/ncu_gelu.log 2>&1
tail -3 gpurun_out/ncu_gelu.log
