#!/bin/bash
# round-2 GPU call Z: final build -- tests, smoke, bench, launch list of a step, cfg5 sweep, cfg3 training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/z_*
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest_gpu.log; tail -n 3 gpurun_out/z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/z_smoke.log 2>&1; tail -n 2 gpurun_out/z_smoke.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/z_launches.csv python scripts/profile_step.py > gpurun_out/z_prof.log 2>&1; tail -n 1 gpurun_out/z_prof.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?" >> gpurun_out/z_bench.err; head -c 250 gpurun_out/z_bench.json; echo
timeout 900 python bench.py --sweep > gpurun_out/z_sweep.json 2> gpurun_out/z_sweep.err; head -c 200 gpurun_out/z_sweep.json; echo
timeout 600 python bench.py --train --batch 8 --steps 3 --warmup 1 > gpurun_out/z_train.json 2> gpurun_out/z_train.err; head -c 300 gpurun_out/z_train.json; echo
