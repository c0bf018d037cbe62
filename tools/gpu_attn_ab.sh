#!/bin/bash
# Attention A/B on one box: parity tests (under a timeout: a protocol bug hangs), micro-benchmark of the in-tree library
# against tools/_build/libhriemo_<name>.so for every name given, pipeline trace of the in-tree kernel.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -3
echo "== in-tree"; timeout 200 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | cut -c1-160
for v in "$@"; do
  echo "== $v"; HRIEMO_LIB_PATH=tools/_build/libhriemo_$v.so timeout 200 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | grep -v head_pairs | cut -c1-160
done
timeout 120 python tools/attn_trace.py 64 8 500 500 96 > gpurun_out/trace_ab_500.txt 2>&1; echo "trace exit=$?"
