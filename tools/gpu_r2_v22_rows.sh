#!/bin/bash
# v22: row-kernel timings (pooling with the statistics / PAD bytes staged), then the ncu --set full captures.
mkdir -p gpurun_out
timeout 300 python tools/bench_decoder_gate.py > gpurun_out/v22_dg.log 2>&1; grep "ln_masked_mean\|gate_blend" gpurun_out/v22_dg.log | cut -c1-200
bash tools/gpu_r2_v22_ncu.sh
