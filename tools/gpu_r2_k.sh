#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_input_pipeline.py tests/test_gpu_model.py tests/test_gpu_optimizer.py tests/test_gpu_reference_parity.py tests/test_gpu_simple_kernels.py
CMD="python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline"
$CMD > gpurun_out/bench_r2_k.log 2>&1
grep '^{' gpurun_out/bench_r2_k.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'fp32', round(d['e2e']['fp32_input']['value'],1), d['clocks']['sm_mhz'], 'gemm', round(d['roofline']['achieved'],1), d['roofline']['frac'], d['gpu_launches'])"
CMD2="python bench.py --steps 1 --warmup 3 --no-e2e --no-gpu-reference --no-cpu-baseline --no-roofline-leg"
$CMD2 > gpurun_out/plain_r2_k.log 2>&1 && \
ncu --set full --clock-control none -k regex:"target_mse_kernel|gather_tubes_kernel|tube_mask_kernel|adamw_kernel|assemble_fwd_kernel|loss_finish_kernel|zero_rows" -c 14 -o gpurun_out/r02_prof_hbm3 $CMD2 > gpurun_out/ncu_r2_k.log 2>&1
tail -2 gpurun_out/ncu_r2_k.log
