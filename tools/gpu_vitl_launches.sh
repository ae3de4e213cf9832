#!/bin/bash
# launch list of one ViT-L (configs[3], B = 16) step
mkdir -p gpurun_out
CMD="python bench.py --model pretrain_videomae_large_patch16_224 --batch 16 --steps 1 --warmup 3 --no-e2e --no-gpu-reference --no-cpu-baseline --no-roofline-leg"
$CMD > gpurun_out/plain_vitl.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_vitl_launches.csv $CMD > gpurun_out/ncu_vitl.log 2>&1
python tools/launch_summary.py gpurun_out/r02_vitl_launches.csv > gpurun_out/r02_vitl_launches_summary.txt 2>&1
head -40 gpurun_out/r02_vitl_launches_summary.txt
