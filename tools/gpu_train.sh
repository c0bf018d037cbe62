#!/bin/bash
# Training-step records: every GPU test, smoke(), the training-step benchmark at two batch sizes, and one
# ncu --set full capture of the attention-backward kernels (first 4 launches of a small step).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for b in 128 512; do
  timeout 120 python tools/bench_train.py --batch $b --steps 3 --warmup 2 > gpurun_out/bench_train_b$b.json 2> gpurun_out/bench_train_b$b.err || tail -3 gpurun_out/bench_train_b$b.err
  cat gpurun_out/bench_train_b$b.json
done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_d -c 4 -f -o gpurun_out/prof_attn_bwd \
  python tools/bench_train.py --batch 32 --steps 1 --warmup 1 > gpurun_out/ncu_attn_bwd.log 2>&1
tail -2 gpurun_out/ncu_attn_bwd.log
