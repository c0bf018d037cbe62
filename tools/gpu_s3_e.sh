#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
echo "== gate pooling new"; timeout 120 python tools/bench_lmm.py
echo "== gate pooling base"; HRIEMO_LIB_PATH=tools/_build/libhriemo_base.so timeout 120 python tools/bench_lmm.py
timeout 600 python tools/bench_ragged.py 2>&1 | tail -1 | tee gpurun_out/bench_ragged.json
timeout 600 python bench.py > gpurun_out/bench_s3e.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench_s3e.log | cut -c1-400
