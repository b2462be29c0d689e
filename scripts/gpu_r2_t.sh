#!/bin/bash
# round-2 GPU call T: ncu --set full of the GroupNorm apply / finalize kernels and of the small-K GEMMs with 3-4 epilogue warpgroups
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/t_*
timeout 200 python scripts/gn_one.py > gpurun_out/t_gn_one.log 2>&1; tail -n 1 gpurun_out/t_gn_one.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_apply_ms -s 3 -c 1 -o gpurun_out/t_prof_gn_apply \
    python scripts/gn_one.py > gpurun_out/t_prof_gn_apply.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_finalize -s 3 -c 1 -o gpurun_out/t_prof_gn_fin \
    python scripts/gn_one.py > gpurun_out/t_prof_gn_fin.log 2>&1
for c in qkv geglu res; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s 3 -c 1 -o gpurun_out/t_prof_gemm_$c \
      python scripts/one_gemm.py $c > gpurun_out/t_prof_gemm_$c.log 2>&1
done
ls -la gpurun_out/t_*.ncu-rep
