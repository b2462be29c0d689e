#!/bin/bash
# round-2 GPU call B: fused LN / GN-statistics epilogues (tests + step bench + launch list), attention variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( cd storygen_b200/csrc && timeout 300 ./selftest ) > gpurun_out/b_selftest.log 2>&1
echo "selftest rc=$?" >> gpurun_out/b_selftest.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/b_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest_gpu.log
for v in default nodefer p49d pph1 pph2; do
  if [ $v = default ]; then lib=""; else lib="$PWD/storygen_b200/csrc/alt/$v/libstorygen_b200.so"; fi
  echo "== variant $v" >> gpurun_out/b_attn_bench.log
  STORYGEN_B200_LIB=$lib timeout 300 python scripts/attn_bench.py 2>&1 | grep -E "d= 40" >> gpurun_out/b_attn_bench.log
done
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench rc=$?" >> gpurun_out/b_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/b_launches.csv python scripts/profile_step.py > gpurun_out/b_prof.log 2>&1
tail -3 gpurun_out/b_selftest.log; tail -4 gpurun_out/b_pytest_gpu.log; cat gpurun_out/b_attn_bench.log; head -c 300 gpurun_out/b_bench.json; tail -2 gpurun_out/b_bench.err
timeout 300 python scripts/power_probe.py > gpurun_out/b_power.log 2>&1
cat gpurun_out/b_power.log
