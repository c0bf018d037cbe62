#!/bin/bash
# Closing records of round 2 (v24) on one box: every GPU test, smoke(), the default bench line, the reference arm.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/v24_gpu_tests.txt 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/v24_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/v24_bench_default.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/v24_bench_default.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v24_bench_ref.log 2>&1; echo "ref exit=$?"; tail -1 gpurun_out/v24_bench_ref.log | cut -c1-200
