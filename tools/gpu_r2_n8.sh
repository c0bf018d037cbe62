#!/bin/bash
# 8-GPU records: concurrent H2D probe, the north-star bench (incl. e2e and the training step with its all-reduce),
# config 3 (MOSEI wrapper) and the longest config-4 point, each one JSON line.
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
nproc > gpurun_out/r2_n${N}_host.txt; nvidia-smi topo -m >> gpurun_out/r2_n${N}_host.txt 2>&1
timeout 300 $TR tools/h2d_concurrent_probe.py > gpurun_out/r2_h2d_probe_n$N.log 2>&1; echo "probe exit=$?"; tail -1 gpurun_out/r2_h2d_probe_n$N.log | cut -c1-700
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_ns_n$N.log 2>&1; echo "ns exit=$?"; tail -1 gpurun_out/r2_bench_ns_n$N.log | cut -c1-300
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --workload cfg3 --no-train > gpurun_out/r2_bench_cfg3_n$N.log 2>&1; echo "cfg3 exit=$?"; tail -1 gpurun_out/r2_bench_cfg3_n$N.log | cut -c1-300
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --workload long --no-train --no-e2e > gpurun_out/r2_bench_long_n$N.log 2>&1; echo "long exit=$?"; tail -1 gpurun_out/r2_bench_long_n$N.log | cut -c1-300
