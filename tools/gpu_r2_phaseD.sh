#!/bin/bash
# Round-2 phase D: training-step fusions (ReLU mask in the dX GEMM epilogue, warp-per-row D kernel, one-round LayerNorm
# backward): backward / ops tests, then the training-step benchmark (CUDA graph and eager breakdown) at B = 512.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_ops_gpu.py tests/test_train_boundary_gpu.py -m gpu -x -q > gpurun_out/r2d_tests.txt 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/r2d_tests.txt
for cfg in "512 --graph" "512" "128 --graph"; do
  set -- $cfg
  timeout 180 python tools/bench_train.py --batch $1 $2 --steps 5 --warmup 4 > "gpurun_out/r2d_bench_train_b$1$2.json" 2> gpurun_out/bt.err || tail -3 gpurun_out/bt.err
  python -c "import json;d=json.load(open('gpurun_out/r2d_bench_train_b$1$2.json'));print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2), round(v['tflops'])) for k,v in d.get('breakdown',{}).items()})"
done
