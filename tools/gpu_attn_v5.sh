#!/bin/bash
# Attention v5 softmax loop: parity tests, then the micro-benchmark of the in-tree library against the variants named on
# the command line (tools/_build/libhriemo_<name>.so), then the pipeline trace of the in-tree kernel (trace build).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -3
echo "== in-tree"; timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | cut -c1-160
for v in "$@"; do
  echo "== $v"; HRIEMO_LIB_PATH=tools/_build/libhriemo_$v.so timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | grep -v head_pairs | cut -c1-160
done
timeout 300 python tools/attn_trace.py 64 8 500 500 96 > gpurun_out/trace_v5_500.txt 2>&1; echo "trace exit=$?"
