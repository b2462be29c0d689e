#!/bin/bash
# round-2 GPU call Y: attn_text_kernel (77-token cross-attention, four CTAs per SM) against the two-CTA SHORT path
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/y_*
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -m gpu -q -x -k "attention" > gpurun_out/y_pytest_attn.log 2>&1; tail -n 3 gpurun_out/y_pytest_attn.log
for v in text0 default; do
  t=1; [ $v = text0 ] && t=0
  echo "== $v" >> gpurun_out/y_attn_bench.log
  SG_ATTN_TEXT=$t timeout 300 python scripts/attn_bench.py 2>&1 | grep "nk=    77" >> gpurun_out/y_attn_bench.log
done
cat gpurun_out/y_attn_bench.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/y_pytest_all.log 2>&1; tail -n 3 gpurun_out/y_pytest_all.log
for v in text0 default text0 default; do
  t=1; [ $v = text0 ] && t=0
  SG_ATTN_TEXT=$t timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/y_bench_$v.json 2> gpurun_out/y_bench_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/y_bench_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"],4), round(d["ms_per_step"]/50,3), d["clocks"]["sm_mhz"], d.get("parity",{}).get("max_abs_delta_step2"))
P
done
