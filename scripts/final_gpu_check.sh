#!/bin/bash
# One GPU-box pass that produces everything the round summary cites (run under gpurun from the repo root).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
python bench.py 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json; cut -c1-300 gpurun_out/final_bench.json
python scripts/gemm_bench.py > gpurun_out/final_gemm_bench.log 2>&1; tail -16 gpurun_out/final_gemm_bench.log
python scripts/attn_bench.py > gpurun_out/final_attn_bench.log 2>&1; cat gpurun_out/final_attn_bench.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv \
    python scripts/profile_step.py > gpurun_out/final_prof.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel.*160 -s 1 -c 1 \
    -o gpurun_out/final_prof_conv python scripts/profile_step.py > gpurun_out/final_prof_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 2 -c 1 \
    -o gpurun_out/final_prof_attn40 python scripts/attn_one.py > gpurun_out/final_prof_attn40.log 2>&1
ls -la gpurun_out/*.ncu-rep
