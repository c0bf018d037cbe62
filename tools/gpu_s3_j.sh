#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_bucketing_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python bench.py --no-cpu > gpurun_out/bench_s3j.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench_s3j.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
