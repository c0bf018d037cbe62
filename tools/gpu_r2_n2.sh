#!/bin/bash
# 2-GPU record of the closing build: the north-star bench under torchrun (incl. e2e and the training step with its all-reduce)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543"
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_ns_n2.log 2>&1; echo "ns exit=$?"; tail -1 gpurun_out/r2g_bench_ns_n2.log | cut -c1-300
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2g_bench_ref_n2.log 2>&1; echo "ref exit=$?"; tail -1 gpurun_out/r2g_bench_ref_n2.log | cut -c1-200
