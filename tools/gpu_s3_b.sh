#!/bin/bash
# Session-3 call B: bucketing tests, host cast probe, e2e timelines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bucketing_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -15
timeout 300 python tools/host_cast_probe.py 2>&1 | tee gpurun_out/host_cast_probe.txt
(timeout 300 python tools/e2e_timeline.py; timeout 300 python tools/e2e_timeline.py --slab 256; timeout 300 python tools/e2e_timeline.py --slab 256 --every 1; timeout 300 python tools/e2e_timeline.py --ragged; timeout 300 python tools/e2e_timeline.py --ragged --bucket) 2>&1 | tee gpurun_out/e2e_timeline_s3.txt
