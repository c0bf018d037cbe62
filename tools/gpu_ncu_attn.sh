#!/bin/bash
# ncu --set full on attention launches of one forward at B=1024 (after 3 warm-up forwards = 24 attention launches):
# $1 = launches to skip (24 = audio self-attention 500x500 of layer 0), $2 = count
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --batch 1024"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_fwd3 -s ${1:-24} -c ${2:-1} -f -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_attn.log | cut -c1-200
