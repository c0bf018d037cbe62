#!/bin/bash
# CTA-pair wgrad against the single-CTA form (HRIEMO_WGRAD_PAIR=0), one build, one box; tests first, under a timeout.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_backward_gpu.py tests/test_abi.py tests/test_fuzz_gpu.py -q -m gpu -x -k "wgrad or linear_backward or trainer or train or guard" 2>&1 | tail -3
for v in 1 0 1 0; do echo "== HRIEMO_WGRAD_PAIR=$v"; HRIEMO_WGRAD_PAIR=$v timeout 200 python tools/bench_kernels.py --only wgrad 2>&1 | cut -c1-215; done
