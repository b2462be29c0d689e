#!/bin/bash
# round-2 GPU call K: attention one-time warpgroup skew variants (+ timeline), inference-flow test
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( cd storygen_b200/csrc && timeout 300 ./selftest attn ) > gpurun_out/k_selftest_attn.log 2>&1; tail -1 gpurun_out/k_selftest_attn.log
for v in default skew0 skew3 skew4 skew2p49; do
  if [ $v = default ]; then lib=""; else lib="$PWD/storygen_b200/csrc/alt/$v/libstorygen_b200.so"; fi
  echo "== variant $v" >> gpurun_out/k_attn_bench.log
  STORYGEN_B200_LIB=$lib timeout 300 python scripts/attn_bench.py 2>&1 | grep -E "d= 40 nq= 4096 nk= ( 4096|12288)" >> gpurun_out/k_attn_bench.log
done
cat gpurun_out/k_attn_bench.log
( cd storygen_b200/csrc && timeout 120 ./selftest timeline ) 2>&1 | grep -v gtimeline > gpurun_out/k_timeline.log; head -20 gpurun_out/k_timeline.log
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q > gpurun_out/k_pytest.log 2>&1; tail -3 gpurun_out/k_pytest.log
