#!/bin/bash
# round-2 GPU call L: attention explicit software-pipelined exp loop variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( cd storygen_b200/csrc && timeout 300 ./selftest attn ) > gpurun_out/l_selftest_attn.log 2>&1; tail -1 gpurun_out/l_selftest_attn.log
for v in default ord ordp49 ordp01 ordp00; do
  if [ $v = default ]; then lib=""; else lib="$PWD/storygen_b200/csrc/alt/$v/libstorygen_b200.so"; fi
  echo "== variant $v" >> gpurun_out/l_attn_bench.log
  STORYGEN_B200_LIB=$lib timeout 300 python scripts/attn_bench.py 2>&1 | grep -E "d= 40 nq= 4096 nk= ( 4096|12288)" >> gpurun_out/l_attn_bench.log
done
cat gpurun_out/l_attn_bench.log
( cd storygen_b200/csrc && LD_LIBRARY_PATH=$PWD/alt/ord timeout 120 ./selftest attn ) 2>&1 | tail -1
( cd storygen_b200/csrc && timeout 120 ./selftest timeline ) 2>&1 | grep -v gtimeline > gpurun_out/l_timeline.log; head -9 gpurun_out/l_timeline.log
