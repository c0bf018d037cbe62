#!/bin/bash
# v20 (streaming gate blend, warp-private decoder attention): kernel A / B, every GPU test, smoke(), the default bench line.
mkdir -p gpurun_out
timeout 300 python tools/bench_decoder_gate.py > gpurun_out/v20_dg_v2.log 2>&1; echo "dg exit=$?"; grep -c kernel gpurun_out/v20_dg_v2.log
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/v20_gpu_tests.txt 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/v20_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/v20_bench_default.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/v20_bench_default.log | cut -c1-400
