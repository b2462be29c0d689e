#!/bin/bash
# round-2 GPU call M: measurements of the committed state (tests, smoke, bench arms, cfg4 batch, launch list, ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest_gpu.log; tail -3 gpurun_out/m_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/m_smoke.log 2>&1; tail -2 gpurun_out/m_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?" >> gpurun_out/m_bench.err; head -c 250 gpurun_out/m_bench.json; echo
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/m_bench_ref.json 2> gpurun_out/m_bench_ref.err; echo "ref rc=$?" >> gpurun_out/m_bench_ref.err; head -c 300 gpurun_out/m_bench_ref.json; echo
timeout 900 python bench.py --batch 8 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/m_bench_b8.json 2> gpurun_out/m_bench_b8.err; head -c 250 gpurun_out/m_bench_b8.json; echo
timeout 300 python scripts/attn_bench.py > gpurun_out/m_attn_bench.log 2>&1
timeout 300 python scripts/gemm_bench.py > gpurun_out/m_gemm_bench.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/m_launches.csv python scripts/profile_step.py > gpurun_out/m_prof.log 2>&1; tail -1 gpurun_out/m_prof.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 3 -c 1 -o gpurun_out/m_prof_attn40 \
    python scripts/attn_one.py > gpurun_out/m_prof_attn40.log 2>&1; ls -la gpurun_out/m_prof_attn40.ncu-rep
