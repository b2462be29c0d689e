#!/bin/bash
# round-2 GPU call R: full GPU suite of the current build; bench with and without the feature pass' dead tail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r_*
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r_pytest_all.log 2>&1; tail -n 3 gpurun_out/r_pytest_all.log
for v in tail default tail default; do
  t=0; [ $v = tail ] && t=1
  SG_FEATURE_TAIL=$t timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r_bench_$v.json 2> gpurun_out/r_bench_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r_bench_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"],4), round(d["ms_per_step"]/50,3), d["clocks"]["sm_mhz"], d.get("gpu_launches"))
P
done
