#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_gemm.py tests/test_gpu_model.py tests/test_gpu_finetune.py tests/test_gpu_reference_parity.py tests/test_gpu_optimizer.py
grep -h "after 6 steps" gpurun_out/test_gpu_finetune.log
N=20 python tools/prof_gemm.py 2>&1 | grep -v wgrad | head -12
python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_r2_m.log 2>&1
grep '^{' gpurun_out/bench_r2_m.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['achieved'],1), d['gpu_launches'])"
