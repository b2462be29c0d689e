#!/bin/bash
# round-2 GPU call I: attention P-release barriers, train bench (cfg3), sweep (cfg5), full test suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( cd storygen_b200/csrc && timeout 300 ./selftest attn ) > gpurun_out/i_selftest_attn.log 2>&1; tail -1 gpurun_out/i_selftest_attn.log
timeout 300 python scripts/attn_bench.py > gpurun_out/i_attn_bench.log 2>&1; cat gpurun_out/i_attn_bench.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
echo "bench rc=$?" >> gpurun_out/i_bench.err; head -c 250 gpurun_out/i_bench.json; echo; tail -1 gpurun_out/i_bench.err
timeout 900 python bench.py --train --batch 8 --steps 3 --warmup 1 > gpurun_out/i_bench_train.json 2> gpurun_out/i_bench_train.err
echo "train rc=$?" >> gpurun_out/i_bench_train.err; head -c 400 gpurun_out/i_bench_train.json; echo; tail -3 gpurun_out/i_bench_train.err
timeout 900 python bench.py --sweep > gpurun_out/i_bench_sweep.json 2> gpurun_out/i_bench_sweep.err
echo "sweep rc=$?" >> gpurun_out/i_bench_sweep.err; head -c 600 gpurun_out/i_bench_sweep.json; echo; tail -2 gpurun_out/i_bench_sweep.err
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/i_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_pytest_gpu.log; tail -4 gpurun_out/i_pytest_gpu.log
