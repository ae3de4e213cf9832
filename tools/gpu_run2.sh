mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_gemm.py | tail -3
for i in 1 2; do
echo "--- new"; N=20 python tools/prof_gemm.py | grep -v wgrad | grep -v ln_
echo "--- v1 (old)"; MOFO_B200_LIB=tools/variants/libmofo_gelu_v1.so ONLY=gelu N=20 python tools/prof_gemm.py
done
