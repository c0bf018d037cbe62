#!/bin/bash
# 8-GPU record of the closing build (v22): the north-star bench under torchrun (forward, e2e, training step with its all-reduce)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/v22_bench_ns_n8.log 2>&1; echo "ns exit=$?"; tail -1 gpurun_out/v22_bench_ns_n8.log | cut -c1-300
