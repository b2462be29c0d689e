#!/bin/bash
# round-2 GPU call W (2 GPUs): the bench as the driver launches it at N=2, both arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/w_*
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/w_bench_2gpu.json 2> gpurun_out/w_bench_2gpu.err; echo "rc=$?"; head -c 300 gpurun_out/w_bench_2gpu.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/w_bench_ref_2gpu.json 2> gpurun_out/w_bench_ref_2gpu.err; echo "rc=$?"; head -c 300 gpurun_out/w_bench_ref_2gpu.json; echo
