#!/bin/bash
# same-call A/B of one environment switch:  tools/gpu_ab.sh VAR valueA valueB
mkdir -p gpurun_out
for v in $2 $3 $2 $3; do
  env $1=$v python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-e2e --no-roofline-leg > gpurun_out/bench_ab_$1_$v.log 2>&1
  grep '^{' gpurun_out/bench_ab_$1_$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1=$v', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['gpu_launches'])"
done
