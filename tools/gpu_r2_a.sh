#!/bin/bash
# round-2 first GPU pass: all GPU tests, 1-GPU bench (with the reference's own GPU step), staged-optimizer A/B, launch list
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_*.py
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_a.log 2> gpurun_out/bench_r2_a.err
tail -c 3000 gpurun_out/bench_r2_a.log
MOFO_STAGED_OPT=0 python bench.py --steps 20 --warmup 5 --no-e2e --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_r2_a_nostaged.log 2>&1
tail -c 600 gpurun_out/bench_r2_a_nostaged.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2_a.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-gpu-reference --no-cpu-baseline > gpurun_out/ncu_r2_a.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2_a.csv > gpurun_out/launches_r2_a_summary.txt 2>&1
head -40 gpurun_out/launches_r2_a_summary.txt
