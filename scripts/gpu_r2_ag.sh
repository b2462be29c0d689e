#!/bin/bash
# round-2 GPU call AG: compute-sanitizer memcheck over the tiny-UNet tests (engine paths: folded LayerNorm, GroupNorm partials,
# sub-pixel up-sampling, text attention kernel, feature pass cut at its last tap, fused q2|q3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ag_*
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_unet.py -m gpu -q -x \
    -k "golden or resolutions or batch_invariance or text_kv or feature_pass or denoise_loop" > gpurun_out/ag_memcheck_unet.log 2>&1; echo "memcheck rc=$?"; tail -n 6 gpurun_out/ag_memcheck_unet.log
