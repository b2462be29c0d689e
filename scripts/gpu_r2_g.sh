#!/bin/bash
# round-2 GPU call G: ncu --set full captures of the attention core and of three GEMM classes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in qkv geglu res conv; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s 3 -c 1 -o gpurun_out/g_prof_gemm_$c \
      python scripts/one_gemm.py $c > gpurun_out/g_prof_gemm_$c.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 3 -c 1 -o gpurun_out/g_prof_attn40 \
    python scripts/attn_one.py > gpurun_out/g_prof_attn40.log 2>&1
ls -la gpurun_out/g_prof_*.ncu-rep
