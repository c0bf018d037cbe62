#!/bin/bash
# v21 (LayerNorm backward on cp.async row rings, staged decoder attention backward, parallel LN parameter-gradient sums):
# tests of the touched kernels, kernel A / B, the training step both ways on one box.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_backward_gpu.py tests/test_dropout_gpu.py tests/test_fuzz_gpu.py -m gpu -x -q > gpurun_out/v21_tests.txt 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/v21_tests.txt
timeout 300 python tools/bench_decoder_gate.py > gpurun_out/v21_dg_v2.log 2>&1; grep "layernorm_backward\|attention_backward" gpurun_out/v21_dg_v2.log | cut -c1-200
HRIEMO_LN_BWD_V1=1 HRIEMO_DECODER_ATTN_BWD_V1=1 timeout 300 python tools/bench_decoder_gate.py > gpurun_out/v21_dg_v1.log 2>&1; grep "layernorm_backward\|attention_backward" gpurun_out/v21_dg_v1.log | cut -c1-200
for i in 1 2; do
  HRIEMO_LN_BWD_V1=1 HRIEMO_DECODER_ATTN_BWD_V1=1 timeout 200 python tools/bench_train.py --batch 512 --graph --steps 5 --warmup 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('former', round(d['value'],1), round(d['ms_per_step'],2))"
  timeout 200 python tools/bench_train.py --batch 512 --graph --steps 5 --warmup 4 2>/dev/null | tee gpurun_out/v21_bench_train_b512_graph.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v21   ', round(d['value'],1), round(d['ms_per_step'],2))"
done
