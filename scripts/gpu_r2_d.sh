#!/bin/bash
# round-2 GPU call C: TMA-store epilogue (tests, step bench, launch list), attention tail reorder
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( cd storygen_b200/csrc && timeout 300 ./selftest ) > gpurun_out/d_selftest.log 2>&1
echo "selftest rc=$?" >> gpurun_out/d_selftest.log
grep -E "FAIL|SELFTEST" gpurun_out/d_selftest.log | head -20
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/d_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest_gpu.log
tail -8 gpurun_out/d_pytest_gpu.log

timeout 300 python scripts/gemm_bench.py > gpurun_out/d_gemm_bench.log 2>&1; tail -30 gpurun_out/d_gemm_bench.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
echo "bench rc=$?" >> gpurun_out/d_bench.err
head -c 300 gpurun_out/d_bench.json; tail -2 gpurun_out/d_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/d_launches.csv python scripts/profile_step.py > gpurun_out/d_prof.log 2>&1
tail -2 gpurun_out/d_prof.log
