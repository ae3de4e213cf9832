#!/bin/bash
# quick 2-GPU check: DP parity test (1e-5) and one bench line at N = 2 (dp_check inside)
mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_dp.py
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_r2_2gpu_s3.log 2>&1
grep '^{' gpurun_out/bench_r2_2gpu_s3.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2gpu', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['dp_check'])"
