mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_attention.py | tail -3
for i in 1 2 3; do
echo "--- pair barrier"; PB=32 python tools/prof_attn.py | tail -1
echo "--- cta barrier"; MOFO_B200_LIB=tools/variants/libmofo_nopair.so PB=32 python tools/prof_attn.py | tail -1
done
