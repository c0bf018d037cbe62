#!/bin/bash
# attention micro-benchmark of the in-tree library and of every tools/_build/libhriemo_<name>.so named on the command line
echo "== in-tree"; timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | cut -c1-150
for v in "$@"; do
  echo "== $v"; HRIEMO_LIB_PATH=tools/_build/libhriemo_$v.so timeout 300 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | cut -c1-150
done
