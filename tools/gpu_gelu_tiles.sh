#!/bin/bash
# does the tile choice for the two GELU GEMM shapes still hold after the epilogue got cheaper?  (prof_gemm.py cases)
for cfg in "" "MOFO_FORCE_BN=256" "MOFO_FORCE_BN=224" "MOFO_FORCE_BN=192" "MOFO_FORCE_BN=128" "MOFO_GEMM_2CTA=0 MOFO_FORCE_BN=256" "MOFO_GEMM_2CTA=0 MOFO_FORCE_BN=192" "MOFO_GEMM_2CTA=0 MOFO_FORCE_BN=128"; do
  echo "== ${cfg:-cost model}"; env $cfg ONLY=gelu N=20 python tools/prof_gemm.py | grep fc1
done
