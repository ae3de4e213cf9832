#!/bin/bash
# final evidence pass (1 GPU): smoke, every GPU test, the default bench (both arms), launch list of the last step
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_r2.log 2>&1; tail -1 gpurun_out/smoke_r2.log
bash tools/gpu_checks.sh tests/test_gpu_*.py
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_final_ref.log 2>&1; grep '^{' gpurun_out/bench_r2_final_ref.log | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.log 2> gpurun_out/bench_r2_final.err
grep '^{' gpurun_out/bench_r2_final.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'fp32', round(d['e2e']['fp32_input']['value'],1), d['clocks'], 'gemm', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), round(d['roofline']['frac_of_burst_peak'],3), 'ref', {k:round(v['value'],1) for k,v in d['gpu_reference'].items() if isinstance(v,dict)}, 'cpu', round(d['cpu_baseline']['value'],2), d['gpu_launches'])"
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-gpu-reference --no-cpu-baseline --no-roofline-leg"
$CMD > gpurun_out/plain_r2_final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_r2_final.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
head -12 gpurun_out/r02_launches_summary.txt
