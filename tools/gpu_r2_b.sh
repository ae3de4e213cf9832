#!/bin/bash
# round-2 second GPU pass: single-pass attention kernels (tests + timing), reference parity with per-tensor report
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_attention.py tests/test_gpu_simple_kernels.py tests/test_gpu_reference_parity.py tests/test_gpu_model.py
grep -h "sum_i dQ_i" gpurun_out/test_gpu_attention.log
for small in 1 0; do
  echo "== encoder attention shape, MOFO_ATTN_SMALL=$small"
  MOFO_ATTN_SMALL=$small PB=32 PS=160 PH=12 timeout 120 python tools/prof_attn.py 2>&1 | tail -2
done
python bench.py --steps 20 --warmup 5 --no-e2e --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_r2_b.log 2>&1
tail -c 700 gpurun_out/bench_r2_b.log
MOFO_ATTN_SMALL=0 python bench.py --steps 20 --warmup 5 --no-e2e --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_r2_b_stream.log 2>&1
tail -c 700 gpurun_out/bench_r2_b_stream.log
