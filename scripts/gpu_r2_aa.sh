#!/bin/bash
# round-2 GPU call AA (2 GPUs): data-parallel stage-2 training step (cfg3), gradient all-reduce over NCCL
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/aa_*
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --train --gpus 2 --batch 8 --steps 3 --warmup 1 > gpurun_out/aa_train_2gpu.json 2> gpurun_out/aa_train_2gpu.err; echo "rc=$?"; tail -n 3 gpurun_out/aa_train_2gpu.err; head -c 400 gpurun_out/aa_train_2gpu.json; echo
