#!/bin/bash
# 2-GPU pass: DP parity test (1e-5), bench at N=2 with the staged optimizer (default for N>1) and without, dp_check
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_dp.py
PORT=29541
for staged in 1 0; do
  MOFO_STAGED_OPT=$staged python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_r2_2gpu_staged$staged.log 2>&1
  PORT=$((PORT+1))
  grep '^{' gpurun_out/bench_r2_2gpu_staged$staged.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('staged=$staged', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'u8', round(d['e2e']['uint8_input']['value'],1), d['dp_check'])"
done
python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_r2_1gpu_e.log 2>&1
grep '^{' gpurun_out/bench_r2_1gpu_e.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1gpu', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
