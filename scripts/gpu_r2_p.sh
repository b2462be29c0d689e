#!/bin/bash
# round-2 GPU call P: attn_mc_kernel (four small CTAs per SM) against attn_kernel<40, 2>; poly-mask variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/p_*
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -m gpu -q -x -k "attention" > gpurun_out/p_pytest_attn.log 2>&1; tail -n 3 gpurun_out/p_pytest_attn.log
for v in mc0 default mcpoly0 mcpoly5; do
  lib=""; mc=1
  case $v in
    mc0) mc=0;;
    mcpoly0|mcpoly5) lib="$PWD/storygen_b200/csrc/alt/$v/libstorygen_b200.so";;
  esac
  echo "== $v" >> gpurun_out/p_attn_bench.log
  SG_ATTN_MC=$mc STORYGEN_B200_LIB=$lib timeout 300 python scripts/attn_bench.py 2>&1 | grep "d= 40" >> gpurun_out/p_attn_bench.log
done
cat gpurun_out/p_attn_bench.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/p_pytest_all.log 2>&1; tail -n 3 gpurun_out/p_pytest_all.log
for v in mc0 default; do
  mc=1; [ $v = mc0 ] && mc=0
  SG_ATTN_MC=$mc timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/p_bench_$v.json 2> gpurun_out/p_bench_$v.err
  head -c 200 gpurun_out/p_bench_$v.json; echo
done
