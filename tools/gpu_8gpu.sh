#!/bin/bash
# 8-GPU pass, the way the driver launches it
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_r2_8gpu_ref.log 2>&1
grep '^{' gpurun_out/bench_r2_8gpu_ref.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_8gpu.log 2> gpurun_out/bench_r2_8gpu.err
grep '^{' gpurun_out/bench_r2_8gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('8gpu', round(d['value'],1), round(d['ms_per_step'],3), 'e2e(u8)', round(d['e2e']['value'],1), 'fp32', round(d['e2e']['fp32_input']['value'],1), d['clocks'], d['dp_check'], 'ref', {k:(round(v['value'],1) if 'value' in v else v) for k,v in d['gpu_reference'].items() if isinstance(v,dict)})"
tail -5 gpurun_out/bench_r2_8gpu.err
MOFO_STAGED_OPT=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 20 --warmup 5 --no-gpu-reference --no-e2e > gpurun_out/bench_r2_8gpu_nostaged.log 2>&1
grep '^{' gpurun_out/bench_r2_8gpu_nostaged.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('8gpu nostaged', round(d['value'],1), round(d['ms_per_step'],3))"
