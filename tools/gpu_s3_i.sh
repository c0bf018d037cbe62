#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu_full.sh
timeout 600 python bench.py --workload cfg3 --no-cpu > gpurun_out/bench_cfg3.log 2>&1; tail -1 gpurun_out/bench_cfg3.log | cut -c1-300
timeout 600 python bench.py --workload cfg2 --no-cpu > gpurun_out/bench_cfg2.log 2>&1; tail -1 gpurun_out/bench_cfg2.log | cut -c1-300
