#!/bin/bash
# round-2 GPU call F: pair GEGLU / wide-N, attention ncu full capture, step bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/f_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest_gpu.log
tail -4 gpurun_out/f_pytest_gpu.log
timeout 300 python scripts/gemm_bench.py 2>&1 | grep -v conv > gpurun_out/f_gemm_bench.log; cat gpurun_out/f_gemm_bench.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?" >> gpurun_out/f_bench.err
head -c 300 gpurun_out/f_bench.json; tail -2 gpurun_out/f_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 3 -c 1 -o gpurun_out/f_prof_attn40 \
    python scripts/attn_one.py > gpurun_out/f_prof_attn40.log 2>&1
tail -2 gpurun_out/f_prof_attn40.log
ls -la gpurun_out/f_prof_attn40.ncu-rep
