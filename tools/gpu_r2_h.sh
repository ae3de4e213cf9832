#!/bin/bash
# round-2 evidence pass (1 GPU): all GPU tests, default bench, launch list, ncu --set full of the HBM-bound kernels
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_*.py
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_h.log 2> gpurun_out/bench_r2_h.err
grep '^{' gpurun_out/bench_r2_h.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'u8', round(d['e2e']['uint8_input']['value'],1), d['clocks'], 'ref', {k:round(v['value'],1) for k,v in d['gpu_reference'].items() if isinstance(v,dict)}, 'cpu', d['cpu_baseline']['value'])"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-gpu-reference --no-cpu-baseline"
$CMD > gpurun_out/plain_r2_h.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_r2_h.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
head -45 gpurun_out/r02_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:"target_mse_kernel|gather_tubes_kernel|tube_mask_kernel|adamw_kernel|assemble_fwd_kernel|sq_norm_kernel|loss_finish_kernel" -s 14 -c 7 -o gpurun_out/r02_prof_hbm $CMD > gpurun_out/ncu_r2_h2.log 2>&1
tail -2 gpurun_out/ncu_r2_h2.log
