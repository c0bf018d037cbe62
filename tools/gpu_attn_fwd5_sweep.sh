#!/bin/bash
mkdir -p gpurun_out
HRIEMO_ATTN_FWD5=1 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_dropout_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -2
for v in 0 1 0 1; do echo "== HRIEMO_ATTN_FWD5=$v dh=96"; HRIEMO_ATTN_FWD5=$v timeout 200 python tools/attn_sweep.py 96 8; done
