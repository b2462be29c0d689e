#!/bin/bash
# round-2 GPU call U: approximate-reciprocal SiLU + division-free GroupNorm finalize against the previous build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/u_*
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/u_pytest_all.log 2>&1; tail -n 3 gpurun_out/u_pytest_all.log
for v in base default base default; do
  lib=""; [ $v = base ] && lib="$PWD/storygen_b200/csrc/alt/base/libstorygen_b200.so"
  STORYGEN_B200_LIB=$lib timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/u_bench_$v.json 2> gpurun_out/u_bench_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/u_bench_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"],4), round(d["ms_per_step"]/50,3), d["clocks"]["sm_mhz"], d.get("parity",{}).get("max_abs_delta_step2"))
P
done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/u_launches.csv python scripts/profile_step.py > gpurun_out/u_prof.log 2>&1; tail -n 1 gpurun_out/u_prof.log
