#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -n 30 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run model   python -m pytest tests/test_model_gpu.py -q -m gpu
run smoke   python __graft_entry__.py --smoke
run bench_small python bench.py --steps 2 --warmup 3 --batch 512
run bench   python bench.py --steps 3 --warmup 3
run ref     python bench.py --impl reference --steps 2 --warmup 1
cat gpurun_out/summary.txt
