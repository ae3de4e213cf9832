#!/bin/bash
# ncu --set full captures (with source) of the decoder attention kernels at B = 32 and of one grouped weight-gradient launch
mkdir -p gpurun_out
PB=32 python tools/prof_attn.py > gpurun_out/prof_attn_plain.log 2>&1 || exit 1
tail -1 gpurun_out/prof_attn_plain.log
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 4 -o gpurun_out/r02_attn_dec -f \
  env PB=32 python tools/prof_attn.py > gpurun_out/ncu_attn.log 2>&1
tail -1 gpurun_out/ncu_attn.log
