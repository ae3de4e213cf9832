#!/bin/bash
mkdir -p gpurun_out
run() {  # name, extra bench args (quoted), env...
  name=$1; extra=$2; shift; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus 8 --steps 20 --warmup 5 --no-gpu-reference --no-e2e --no-roofline-leg $extra > gpurun_out/bench_r2_8gpu_$name.log 2>&1
  grep '^{' gpurun_out/bench_r2_8gpu_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
}
run default "" A=1
run ctas16 "" NCCL_MAX_CTAS=16
run ctas8 "" NCCL_MAX_CTAS=8
run stages "" MOFO_ENC_STAGES=4,4,3,1
run vitl "--model pretrain_videomae_large_patch16_224 --batch 16" A=1
python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-e2e --no-roofline-leg --no-cpu-baseline > gpurun_out/bench_r2_8box_1gpu.log 2>&1
grep '^{' gpurun_out/bench_r2_8box_1gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1gpu same box', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
