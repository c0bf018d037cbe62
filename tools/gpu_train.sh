#!/bin/bash
# Training-step records: every GPU test, smoke(), the training-step benchmark (eager with the per-kernel breakdown, and
# CUDA-graph replay) at two batch sizes.  The ncu capture of the attention-backward kernels:
#   ncu --set full --clock-control none --import-source on -k regex:attn_bwd_d -c 4 -f -o gpurun_out/prof_attn_bwd \
#     python tools/bench_train.py --batch 32 --steps 1 --warmup 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for cfg in "128" "128 --graph" "512" "512 --graph"; do
  set -- $cfg
  timeout 120 python tools/bench_train.py --batch $1 $2 --steps 5 --warmup 4 > "gpurun_out/bench_train_v15_b$1$2.json" 2> gpurun_out/bt.err || tail -3 gpurun_out/bt.err
  python -c "import json;d=json.load(open('gpurun_out/bench_train_v15_b$1$2.json'));print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2), round(v['tflops'])) for k,v in d['breakdown'].items()})"
done
# ragged batch (trailing-PAD masks, lengths uniform in [T/2, T]): default attention backward vs its first form
for cfg in "--ragged" "--ragged --attn-bwd-impl 2"; do
  timeout 60 python tools/bench_train.py --batch 128 $cfg --steps 3 --warmup 2 2> gpurun_out/bt.err | tee -a gpurun_out/bench_train_ragged.jsonl | cut -c1-120
done
