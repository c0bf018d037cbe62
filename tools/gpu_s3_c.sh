#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bucketing_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -15
(timeout 300 python tools/e2e_timeline.py; timeout 300 python tools/e2e_timeline.py --ragged; timeout 300 python tools/e2e_timeline.py --ragged --bucket; timeout 300 python tools/e2e_timeline.py --ragged --bucket --every 0) 2>&1 | tee gpurun_out/e2e_timeline_s3c.txt
