#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/bench_kernels.py --only attention"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 5 -c 1 -o gpurun_out/prof_attn3 $CMD > gpurun_out/ncu_attn3.log 2>&1
echo "attn exit=$?"; tail -3 gpurun_out/ncu_attn3.log
