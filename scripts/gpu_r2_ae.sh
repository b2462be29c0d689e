#!/bin/bash
# round-2 GPU call AE: ncu --set full of attn_text_kernel and attn_mc_kernel; compute-sanitizer memcheck over the kernel tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ae_*
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_text -s 3 -c 1 -o gpurun_out/ae_prof_attn_text \
    python scripts/attn_text_one.py > gpurun_out/ae_prof_attn_text.log 2>&1
SG_ATTN_MC=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_mc -s 3 -c 1 -o gpurun_out/ae_prof_attn_mc \
    python scripts/attn_one.py > gpurun_out/ae_prof_attn_mc.log 2>&1
ls -la gpurun_out/ae_*.ncu-rep
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x \
    -k "attention or layernorm_fold or geglu or groupnorm or pair" > gpurun_out/ae_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 6 gpurun_out/ae_memcheck.log
