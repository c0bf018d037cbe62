#!/bin/bash
# 8-GPU record of config 3 (MOSEI wrapper, the reference's published configuration) on the closing build
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 --workload cfg3 --no-train > gpurun_out/v24_bench_cfg3_n8.log 2>&1; echo "cfg3 exit=$?"; tail -1 gpurun_out/v24_bench_cfg3_n8.log | cut -c1-300
