#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bucketing_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -15
timeout 600 python tools/bench_ragged.py 2>&1 | tail -3 | tee gpurun_out/bench_ragged.json
(timeout 300 python tools/e2e_timeline.py --no-ramp; timeout 300 python tools/e2e_timeline.py --ragged --bucket) 2>&1 | tee gpurun_out/e2e_timeline_s3d.txt
