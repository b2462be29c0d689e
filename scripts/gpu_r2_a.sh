#!/bin/bash
# round-2 GPU call A: attention rework correctness + variants, full GPU test suite, first bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt
( cd storygen_b200/csrc && timeout 300 ./selftest attn ) > gpurun_out/a_selftest_attn.log 2>&1
echo "selftest rc=$?" >> gpurun_out/a_selftest_attn.log
( cd storygen_b200/csrc && timeout 120 ./selftest timeline ) > gpurun_out/a_timeline.log 2>&1
for v in default p01 p49 p55 pp11; do
  if [ $v = default ]; then lib=""; else lib="$PWD/storygen_b200/csrc/alt/$v/libstorygen_b200.so"; fi
  echo "== variant $v" >> gpurun_out/a_attn_bench.log
  STORYGEN_B200_LIB=$lib timeout 300 python scripts/attn_bench.py >> gpurun_out/a_attn_bench.log 2>&1
done
for v in p49 pp11; do
  echo "== timeline $v" >> gpurun_out/a_timeline.log
  ( cd storygen_b200/csrc && LD_LIBRARY_PATH=$PWD/alt/$v timeout 120 ./selftest timeline ) >> gpurun_out/a_timeline.log 2>&1
done
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/a_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_bench.err
tail -3 gpurun_out/a_selftest_attn.log; tail -3 gpurun_out/a_pytest_gpu.log; cat gpurun_out/a_attn_bench.log | grep -E "variant|nk= 12288|nk=  4096" ; head -c 600 gpurun_out/a_bench.json
