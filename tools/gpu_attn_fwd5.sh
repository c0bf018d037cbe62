#!/bin/bash
# attention_fwd5_kernel (O / l epilogue on its own warpgroup) against the shipped form, one build, one box:
# parity tests under HRIEMO_ATTN_FWD5=1 (with a timeout: a protocol bug hangs), then the micro-benchmark both ways.
mkdir -p gpurun_out
HRIEMO_ATTN_FWD5=1 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_dropout_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -4
echo "== fwd5"; HRIEMO_ATTN_FWD5=1 timeout 200 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | grep -v head_pairs | cut -c1-140
echo "== fwd3"; HRIEMO_ATTN_FWD5=0 timeout 200 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | grep -v head_pairs | cut -c1-140
echo "== fwd5"; HRIEMO_ATTN_FWD5=1 timeout 200 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | grep -v head_pairs | cut -c1-140
