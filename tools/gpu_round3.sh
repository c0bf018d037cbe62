#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n 30 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run ops     python -m pytest tests/test_ops_gpu.py -q -m gpu -x
run model   python -m pytest tests/test_model_gpu.py -q -m gpu
run kbench  python tools/bench_kernels.py
run bench   python bench.py --steps 3 --warmup 3 --no-cpu
cat gpurun_out/summary.txt
