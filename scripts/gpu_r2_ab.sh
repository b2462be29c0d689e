#!/bin/bash
# round-2 GPU call AB: fused [q2 | q3] projection of the denoise pass: tests + bench A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_*
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/ab_pytest_all.log 2>&1; tail -n 3 gpurun_out/ab_pytest_all.log
for v in sep default sep default; do
  f=1; [ $v = sep ] && f=0
  SG_FUSE_Q23=$f timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ab_bench_$v.json 2> gpurun_out/ab_bench_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/ab_bench_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"],4), round(d["ms_per_step"]/50,3), d["clocks"]["sm_mhz"], d.get("gpu_launches"))
P
done
