#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n 25 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run ops     python -m pytest tests/test_ops_gpu.py -q -m gpu -x
run kbench  python tools/bench_kernels.py --only gemm
run kbencha python tools/bench_kernels.py --only attention
run model   python -m pytest tests/test_model_gpu.py -q -m gpu -x
run bench   python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e
cat gpurun_out/summary.txt
