#!/bin/bash
# v23 on 4 GPUs: the bench under torchrun without the CPU / stock-PyTorch legs (forward, e2e, training step with the overlapped exchange)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29549"
timeout 600 $TR bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu --no-torch --no-ragged > gpurun_out/v23_bench_ns_n4.log 2>&1; echo "ns exit=$?"
tail -1 gpurun_out/v23_bench_ns_n4.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); t=d['train_step']; print(d['value'], d['e2e']['value'], {k:t.get(k) for k in ('value','ms_per_step','allreduce_ms','allreduce_share_of_step','allreduce_overlapped')})"
