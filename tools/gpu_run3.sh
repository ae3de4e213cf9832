mkdir -p gpurun_out
bash tools/gpu_checks.sh tests/test_gpu_gemm.py | tail -2
for i in 1 2; do
echo "--- new (prefetch 2)"; N=20 python tools/prof_gemm.py | grep -v wgrad | grep -v ln_
echo "--- prefetch 1"; MOFO_B200_LIB=tools/variants/libmofo_pf1.so N=20 python tools/prof_gemm.py | grep -v wgrad | grep -v ln_
done
for v in "" tools/variants/libmofo_pf1.so "" tools/variants/libmofo_pf1.so; do
  if [ -z "$v" ]; then unset MOFO_B200_LIB; else export MOFO_B200_LIB=$v; fi
  python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-e2e --no-roofline-leg > gpurun_out/bench_ab.log 2>&1
  grep '^{' gpurun_out/bench_ab.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lib=$v', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['gpu_launches'], d.get('final_loss'))"
done
