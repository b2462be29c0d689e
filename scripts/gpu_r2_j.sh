#!/bin/bash
# round-2 GPU call J: attention early barrier probes, ex2 throughput micro-benchmark
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
scripts/ubench/ex2_rate > gpurun_out/j_ex2_rate.log 2>&1; cat gpurun_out/j_ex2_rate.log
( cd storygen_b200/csrc && timeout 300 ./selftest attn ) > gpurun_out/j_selftest_attn.log 2>&1; tail -1 gpurun_out/j_selftest_attn.log
timeout 300 python scripts/attn_bench.py 2>&1 | grep "d= 40" > gpurun_out/j_attn_bench.log; cat gpurun_out/j_attn_bench.log
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_kernels.py -m gpu -q > gpurun_out/j_pytest.log 2>&1; tail -3 gpurun_out/j_pytest.log
