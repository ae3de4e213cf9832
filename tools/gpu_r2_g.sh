#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_gemm.py tests/test_gpu_model.py
for nt in 3 1; do
  echo "== MOFO_WGRAD_NT=$nt"
  MOFO_WGRAD_NT=$nt ONLY=wgrad N=10 python tools/prof_gemm.py 2>&1 | tail -3
  MOFO_WGRAD_NT=$nt python bench.py --steps 20 --warmup 5 --no-e2e --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_r2_g_nt$nt.log 2>&1
  grep '^{' gpurun_out/bench_r2_g_nt$nt.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nt=$nt', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
done
