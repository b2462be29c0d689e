#!/bin/bash
# round-2 GPU call O: 2 / 3 / 4 epilogue warpgroups in the small-K GEMM kernels (SG_GEMM_EG)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/o_*
for eg in 4 3; do
  SG_GEMM_EG=$eg timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/o_pytest_eg$eg.log 2>&1; tail -2 gpurun_out/o_pytest_eg$eg.log
done
for eg in 2 3 4; do
  echo "== eg$eg" >> gpurun_out/o_gemm_bench.log
  SG_GEMM_EG=$eg timeout 300 python scripts/gemm_bench.py 2>&1 | grep -v conv >> gpurun_out/o_gemm_bench.log
  SG_GEMM_EG=$eg timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/o_bench_eg$eg.json 2> gpurun_out/o_bench_eg$eg.err
  head -c 200 gpurun_out/o_bench_eg$eg.json; echo
done
cat gpurun_out/o_gemm_bench.log
