#!/bin/bash
# BASELINE configs[4] (finetune, 1568 tokens) and configs[3] (ViT-L, B = 16) on one GPU, reference on the same GPU beside them
mkdir -p gpurun_out
python bench.py --workload finetune --steps 20 --warmup 5 > gpurun_out/bench_finetune_s3.log 2>&1
grep '^{' gpurun_out/bench_finetune_s3.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('finetune', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v['value'],1) for k,v in d.get('gpu_reference',{}).items() if isinstance(v,dict)})"
python bench.py --model pretrain_videomae_large_patch16_224 --batch 16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vitl_s3.log 2>&1
grep '^{' gpurun_out/bench_vitl_s3.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('vitl', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), {k:round(v['value'],1) for k,v in d.get('gpu_reference',{}).items() if isinstance(v,dict)})"
