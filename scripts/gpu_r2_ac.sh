#!/bin/bash
# round-2 GPU call AC: final build (after the q2|q3 fusion) -- tests, smoke, launch list of a step, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ac_*
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/ac_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ac_pytest_gpu.log; tail -n 3 gpurun_out/ac_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/ac_smoke.log 2>&1; tail -n 2 gpurun_out/ac_smoke.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/ac_launches.csv python scripts/profile_step.py > gpurun_out/ac_prof.log 2>&1; tail -n 1 gpurun_out/ac_prof.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err; echo "bench rc=$?" >> gpurun_out/ac_bench.err; head -c 250 gpurun_out/ac_bench.json; echo
