#!/bin/bash
# Attention A/B on one box: parity tests, micro-benchmark of the in-tree library against tools/_build/libhriemo_base.so,
# pipeline trace of the in-tree kernel (trace build).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -3
echo "== attention new";  timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | cut -c1-200
echo "== attention base"; HRIEMO_LIB_PATH=tools/_build/libhriemo_base.so timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | grep -v head_pairs | cut -c1-200
timeout 300 python tools/attn_trace.py 64 8 500 500 96 > gpurun_out/trace_r2_500.txt 2>&1; echo "trace exit=$?"
