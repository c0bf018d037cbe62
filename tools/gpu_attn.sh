#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n 40 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run attn    python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention and not small"
run kbench_attn python tools/bench_kernels.py --only attention
cat gpurun_out/summary.txt
