#!/bin/bash
# Builds a variant of the library with extra nvcc flags into gpurun_out/variants/<name>.so for A/B runs through
# MOFO_B200_LIB (the product library is untouched):  tools/build_variant.sh trace -DMOFO_ATTN_TRACE
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p mofo_b200/build/variant_$name tools/variants
for s in runtime simple_kernels gemm attention attention_small optimizer motion_kernels; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --use_fast_math "$@" \
    -c mofo_b200/csrc/$s.cu -o mofo_b200/build/variant_$name/$s.o &
done
wait
nvcc -shared -o tools/variants/libmofo_$name.so mofo_b200/build/variant_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo tools/variants/libmofo_$name.so
