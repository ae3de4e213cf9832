#!/bin/bash
# evidence pass 2 (1 GPU): tests touched since, our launch list (last step), the reference's launch list, HBM kernels ncu
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_simple_kernels.py tests/test_gpu_model.py tests/test_gpu_optimizer.py
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-gpu-reference --no-cpu-baseline --no-roofline-leg"
$CMD > gpurun_out/plain_r2_j.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_r2_j.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
head -48 gpurun_out/r02_launches_summary.txt
python tools/ref_step.py bf16 32 3 > gpurun_out/plain_ref_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02_ref_launches.csv python tools/ref_step.py bf16 32 3 > gpurun_out/ncu_ref_step.log 2>&1
cat gpurun_out/plain_ref_step.log | tail -3
python tools/ref_launch_summary.py gpurun_out/r02_ref_launches.csv 3 > gpurun_out/r02_ref_launches_summary.txt 2>&1
head -40 gpurun_out/r02_ref_launches_summary.txt
ncu --set full --clock-control none -k regex:"target_mse_kernel|gather_tubes_kernel|tube_mask_kernel|adamw_kernel|assemble_fwd_kernel|loss_finish_kernel|layernorm_bwd_kernel|attn_small" -s 20 -c 12 -o gpurun_out/r02_prof_hbm2 $CMD > gpurun_out/ncu_r2_j2.log 2>&1
tail -2 gpurun_out/ncu_r2_j2.log
