#!/bin/bash
# Run every self-test case group in its own process (a trap in one case must not hide the others).
cd "$(dirname "$0")/../storygen_b200/csrc" || exit 1
out=../../gpurun_out/selftest_$(date +%H%M%S).log
for f in "gemm" "conv3x3" "1 tile" "d40 self" "d40 text" "kv_map" "fused qkv" "peaked scores" "d80 self" "d80 text" "d160 mid" "d160 self" "long d160" "long d80" "long d40" $EXTRA_FILTERS; do
  timeout 60 ./selftest "$f" 2>&1 | grep -E "PASS|FAIL|error|abort|SELFTEST" | tee -a "$out"
done
