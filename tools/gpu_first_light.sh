#!/bin/bash
# First GPU contact: each test group in its own process with a timeout, so a fault or
# hang in one kernel does not hide the others.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n 25 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run gemm      python -m pytest tests/test_ops_gpu.py -q -m gpu -k "gemm" -x
run attention python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention and not small" 
run elem      python -m pytest tests/test_ops_gpu.py -q -m gpu -k "not gemm and not (attention and not small)"
run bench     python tools/bench_kernels.py
cat gpurun_out/summary.txt
