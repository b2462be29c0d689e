#!/bin/bash
# round-2 GPU call Q: energy per launch of the attention variants (the step is power-capped)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/q_*
for v in mc0 default mcpoly0 mcpoly5 mc0 default; do
  lib=""; mc=1
  case $v in
    mc0) mc=0;;
    mcpoly0|mcpoly5) lib="$PWD/storygen_b200/csrc/alt/$v/libstorygen_b200.so";;
  esac
  VARIANT=$v SG_ATTN_MC=$mc STORYGEN_B200_LIB=$lib timeout 120 python scripts/attn_power.py 2>&1 | tail -n 1 >> gpurun_out/q_attn_power.log
done
cat gpurun_out/q_attn_power.log
