#!/bin/bash
# round-2 GPU call S: energy per launch of attn_kernel<40, 2> variants (polynomial-exp share, ptxas- vs hand-ordered exp loop)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/s_*
for v in default pm00 pm01 pm55 ord0 default; do
  lib=""; [ $v != default ] && lib="$PWD/storygen_b200/csrc/alt/$v/libstorygen_b200.so"
  VARIANT=$v STORYGEN_B200_LIB=$lib timeout 120 python scripts/attn_power.py 2>&1 | tail -n 1 >> gpurun_out/s_attn_power.log
done
cat gpurun_out/s_attn_power.log
