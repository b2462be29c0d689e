#!/bin/bash
# round-2 GPU call AF (4 GPUs): the default bench at N=4 (completes the 1 / 2 / 4 / 8 weak-scaling table)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/af_*
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/af_bench_4gpu.json 2> gpurun_out/af_bench_4gpu.err; echo "rc=$?"; head -c 300 gpurun_out/af_bench_4gpu.json; echo
