#!/bin/bash
# round-2 GPU call AD: asynchronous input staging in _DenoiseLoop.load() (the e2e arm) -- tests + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ad_*
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/ad_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ad_pytest_gpu.log; tail -n 3 gpurun_out/ad_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/ad_bench.json 2> gpurun_out/ad_bench.err; echo "bench rc=$?" >> gpurun_out/ad_bench.err; head -c 250 gpurun_out/ad_bench.json; echo
python - <<P
import json
d=json.loads(open("gpurun_out/ad_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],4), d["e2e"], d["clocks"]["sm_mhz"])
P
