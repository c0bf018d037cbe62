#!/bin/bash
# v23 (overlapped gradient exchange) on 2 GPUs: equality check against the plain exchange, then the bench's training leg.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547"
timeout 600 $TR tools/train_overlap_check.py 512 2>&1 | grep -v "^W\|^\*\*\*" | tee gpurun_out/v23_overlap_check_n2.txt | tail -8
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged > gpurun_out/v23_bench_ns_n2.log 2>&1; echo "ns exit=$?"
tail -1 gpurun_out/v23_bench_ns_n2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); t=d['train_step']; print(d['value'], {k:t.get(k) for k in ('value','ms_per_step','allreduce_ms','allreduce_share_of_step','allreduce_overlapped')})"
