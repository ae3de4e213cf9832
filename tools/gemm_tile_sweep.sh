#!/bin/bash
# tile-configuration sweep for the encoder's M = 5120 GEMM shapes (checks the cost model of pick_tiles)
export SHAPES="5120,2304,768;5120,768,768;5120,3072,768;5120,768,3072;5120,768,2304"
echo "== cost model"; python tools/prof_gemm_shapes.py
for pairs in 0 1; do for bn in 128 192 224 256; do
  if [ $pairs = 0 ] && [ $bn = 224 ]; then continue; fi
  echo "== pairs=$pairs bn=$bn"; MOFO_GEMM_2CTA=$pairs MOFO_FORCE_BN=$bn python tools/prof_gemm_shapes.py
done; done
