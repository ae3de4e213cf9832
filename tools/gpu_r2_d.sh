#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_finetune.py tests/test_gpu_attention.py tests/test_gpu_simple_kernels.py
grep -h "logits rel\|worst gradient" gpurun_out/test_gpu_finetune.log
python bench.py --workload finetune --steps 10 --warmup 3 > gpurun_out/bench_r2_finetune.log 2>&1; tail -c 1500 gpurun_out/bench_r2_finetune.log
python bench.py --model pretrain_videomae_large_patch16_224 --batch 16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_vitl.log 2>&1; tail -c 2500 gpurun_out/bench_r2_vitl.log
python bench.py --steps 20 --warmup 5 --no-e2e --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_r2_d.log 2>&1; tail -c 400 gpurun_out/bench_r2_d.log
