#!/bin/bash
# Session-3 call A: attention epilogue A/B (batched O loads), e2e with the ramped slab plan.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -3
echo "== attention new";  timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | cut -c1-200
echo "== attention base"; HRIEMO_LIB_PATH=tools/_build/libhriemo_base.so timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | cut -c1-200
timeout 600 python bench.py --no-cpu > gpurun_out/bench_s3a.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench_s3a.log | cut -c1-3000
