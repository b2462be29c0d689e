#!/bin/bash
# round-2 GPU call X (8 GPUs): BASELINE cfg4 (batch 8 per GPU, global batch 64) and the default bench at N=8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/x_*
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --batch 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/x_bench_8gpu_cfg4.json 2> gpurun_out/x_bench_8gpu_cfg4.err; echo "rc=$?"; head -c 300 gpurun_out/x_bench_8gpu_cfg4.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/x_bench_8gpu.json 2> gpurun_out/x_bench_8gpu.err; echo "rc=$?"; head -c 300 gpurun_out/x_bench_8gpu.json; echo
