#!/bin/bash
# round-2 GPU call H: training path (attention backward kernel, stage-2 gradient parity)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -s > gpurun_out/h_pytest_train.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest_train.log
grep -E "passed|failed|FAILED|Error|max err|deviat|stage-2" gpurun_out/h_pytest_train.log | head -30
