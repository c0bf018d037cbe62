#!/bin/bash
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -x 2>&1 | tail -3
bash tools/gpu_s3_g.sh
