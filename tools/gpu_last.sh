#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/final_gpu_tests.txt
for cfg in "--ragged" "--ragged --attn-bwd-impl 2"; do
  timeout 40 python tools/bench_train.py --batch 128 $cfg --steps 3 --warmup 2 2> gpurun_out/bt.err | tee -a gpurun_out/bench_train_v16_ragged.jsonl | cut -c1-120
done
