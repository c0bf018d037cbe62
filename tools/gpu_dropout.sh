#!/bin/bash
# Dropout of the training step: its GPU tests, the attention / backward / ops suites (the p = 0 paths must be unchanged),
# the attention micro-benchmark and the training step at p = 0 (neither may have slowed down) and at p = 0.1.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropout_gpu.py -m gpu -x -q > gpurun_out/r2e_dropout_tests.txt 2>&1; echo "dropout exit=$?"; tail -4 gpurun_out/r2e_dropout_tests.txt
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_backward_gpu.py -m gpu -x -q > gpurun_out/r2e_tests.txt 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2e_tests.txt
timeout 200 python tools/bench_kernels.py --only attention 2>&1 | grep -v ragged | grep -v head_pairs | cut -c1-140
for cfg in "512 --graph" "512 --dropout 0.1"; do
  set -- $cfg
  timeout 180 python tools/bench_train.py --batch $1 $2 $3 --steps 5 --warmup 4 > "gpurun_out/r2e_bench_train.json" 2> gpurun_out/bt.err || tail -3 gpurun_out/bt.err
  python -c "import json;d=json.load(open('gpurun_out/r2e_bench_train.json'));print('$cfg', round(d['value'],1), round(d['ms_per_step'],2))"
done
