#!/bin/bash
mkdir -p gpurun_out
HRIEMO_ATTN_FWD4=1 python tools/attn_once.py 512 8 500 500 96 > gpurun_out/fwd4_plain.log 2>&1 && \
HRIEMO_ATTN_FWD4=1 ncu --set full --clock-control none --import-source on -k regex:attention_fwd4 -s 2 -c 1 -f -o gpurun_out/prof_fwd4 python tools/attn_once.py 512 8 500 500 96 > gpurun_out/ncu_fwd4.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/ncu_fwd4.log | cut -c1-200
