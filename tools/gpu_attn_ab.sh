#!/bin/bash
# Same-box A/B of the attention kernel: in-tree library against tools/_build/libhriemo_base.so
echo "== attention new";  timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | cut -c1-170
echo "== attention base"; HRIEMO_LIB_PATH=tools/_build/libhriemo_base.so timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | grep -v head_pairs | cut -c1-170
