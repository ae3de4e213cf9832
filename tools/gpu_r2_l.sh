#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_finetune.py tests/test_gpu_input_pipeline.py tests/test_gpu_simple_kernels.py tests/test_gpu_gemm.py tests/test_gpu_model.py
grep -h "drop_path 0.3" gpurun_out/test_gpu_finetune.log
for cap in 1 2 3 4; do echo "== MOFO_LN_CAP=$cap"; MOFO_LN_CAP=$cap ONLY=ln_bwd N=20 python tools/prof_gemm.py 2>&1 | tail -2; done
python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-e2e > gpurun_out/bench_r2_l.log 2>&1
grep '^{' gpurun_out/bench_r2_l.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['gpu_launches'])"
python bench.py --workload finetune --steps 10 --warmup 3 --no-gpu-reference > gpurun_out/bench_r2_l_ft.log 2>&1; grep '^{' gpurun_out/bench_r2_l_ft.log | cut -c1-260
