#!/bin/bash
# Round-2 closing records on one box: every GPU test, smoke(), the default bench line, the reference arm.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_gpu_tests.txt 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2f_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2f_bench_default.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/r2f_bench_default.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.log 2>&1; echo "ref exit=$?"; tail -1 gpurun_out/r2f_bench_ref.log | cut -c1-300
