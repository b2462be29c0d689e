#!/bin/bash
# round-2 GPU call V: measurements of the final build (tests, smoke, both bench arms, cfg4 per-GPU batch, micro-benches, ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/v_*
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/v_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest_gpu.log; tail -n 3 gpurun_out/v_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/v_smoke.log 2>&1; tail -n 2 gpurun_out/v_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?" >> gpurun_out/v_bench.err; head -c 250 gpurun_out/v_bench.json; echo
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/v_bench_ref.json 2> gpurun_out/v_bench_ref.err; head -c 300 gpurun_out/v_bench_ref.json; echo
timeout 900 python bench.py --batch 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench_b8.json 2> gpurun_out/v_bench_b8.err; head -c 250 gpurun_out/v_bench_b8.json; echo
timeout 300 python scripts/attn_bench.py > gpurun_out/v_attn_bench.log 2>&1
timeout 300 python scripts/gemm_bench.py > gpurun_out/v_gemm_bench.log 2>&1
timeout 300 python scripts/power_probe.py > gpurun_out/v_power_probe.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_apply_ms -s 3 -c 1 -o gpurun_out/v_prof_gn_apply \
    python scripts/gn_one.py > gpurun_out/v_prof_gn_apply.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 3 -c 1 -o gpurun_out/v_prof_attn40 \
    python scripts/attn_one.py > gpurun_out/v_prof_attn40.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s 3 -c 1 -o gpurun_out/v_prof_conv320 \
    python scripts/one_gemm.py conv > gpurun_out/v_prof_conv320.log 2>&1
ls -la gpurun_out/v_*.ncu-rep
