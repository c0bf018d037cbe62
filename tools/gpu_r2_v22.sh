#!/bin/bash
# v22 (attention backward: two elementwise warps per TMEM lane quadrant; gate_stream_grad warp per row; gate MLP layer 1
# of the training forward on the split-operand GEMM): tests, kernel timings, training step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_dropout_gpu.py tests/test_fuzz_gpu.py -m gpu -x -q > gpurun_out/v22_tests.txt 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/v22_tests.txt
timeout 300 python tools/bench_attn_bwd.py 0 2>&1 | tee gpurun_out/v22_attn_bwd.txt | cut -c1-200
for i in 1 2; do
  timeout 200 python tools/bench_train.py --batch 512 --graph --steps 5 --warmup 4 2>/dev/null | tee gpurun_out/v22_bench_train_b512_graph.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v22   ', round(d['value'],1), round(d['ms_per_step'],2))"
done
timeout 200 python tools/bench_train.py --batch 512 --steps 5 --warmup 4 2>/dev/null | tee gpurun_out/v22_bench_train_b512.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('eager', round(d['value'],1), round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2), round(v['tflops'])) for k,v in d['breakdown'].items()})"
