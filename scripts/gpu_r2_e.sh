#!/bin/bash
# round-2 GPU call E: packed-arithmetic epilogue + CTA-pair GEMM (tests, gemm bench A/B, step bench A/B, launch list)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( cd storygen_b200/csrc && timeout 300 ./selftest ) > gpurun_out/e_selftest.log 2>&1
echo "selftest rc=$?" >> gpurun_out/e_selftest.log
grep -E "FAIL|SELFTEST" gpurun_out/e_selftest.log | head -20
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/e_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest_gpu.log
tail -8 gpurun_out/e_pytest_gpu.log
echo "== pair on" > gpurun_out/e_gemm_bench.log
timeout 300 python scripts/gemm_bench.py >> gpurun_out/e_gemm_bench.log 2>&1
echo "== pair off" >> gpurun_out/e_gemm_bench.log
SG_GEMM_PAIR=0 timeout 300 python scripts/gemm_bench.py 2>&1 | grep conv >> gpurun_out/e_gemm_bench.log
cat gpurun_out/e_gemm_bench.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
echo "bench rc=$?" >> gpurun_out/e_bench.err
head -c 300 gpurun_out/e_bench.json; tail -2 gpurun_out/e_bench.err
SG_GEMM_PAIR=0 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/e_bench_nopair.json 2> gpurun_out/e_bench_nopair.err
head -c 300 gpurun_out/e_bench_nopair.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/e_launches.csv python scripts/profile_step.py > gpurun_out/e_prof.log 2>&1
tail -2 gpurun_out/e_prof.log
