#!/bin/bash
# A / B of the streaming gate blend and the warp-private decoder attention against the former kernels.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n 25 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run dg_tests python -m pytest tests/test_ops_gpu.py tests/test_fuzz_gpu.py -q -m gpu -x -k "small_attention or gate or blend or decoder"
run dg_v2 python tools/bench_decoder_gate.py
HRIEMO_GATE_BLEND_V1=1 HRIEMO_DECODER_ATTN_V1=1 run dg_v1 python tools/bench_decoder_gate.py
run dg_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
cat gpurun_out/summary.txt
