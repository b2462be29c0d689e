#!/bin/bash
# round-2 GPU call N: cross-tile prefetch of the epilogue vectors (A/B against the previous build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/n_pytest.log 2>&1; tail -2 gpurun_out/n_pytest.log
for v in default base; do
  if [ $v = default ]; then lib=""; else lib="$PWD/storygen_b200/csrc/alt/$v/libstorygen_b200.so"; fi
  echo "== $v" >> gpurun_out/n_gemm_bench.log
  STORYGEN_B200_LIB=$lib timeout 300 python scripts/gemm_bench.py 2>&1 | grep -v conv >> gpurun_out/n_gemm_bench.log
  STORYGEN_B200_LIB=$lib timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/n_bench_$v.json 2> gpurun_out/n_bench_$v.err
  head -c 200 gpurun_out/n_bench_$v.json; echo
done
cat gpurun_out/n_gemm_bench.log
