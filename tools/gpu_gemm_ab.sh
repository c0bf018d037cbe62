#!/bin/bash
# GEMM A/B on one box: parity tests, micro-benchmark and a short forward bench, in-tree library against tools/_build/libhriemo_base.so
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q -m gpu -x -k "gemm or golden or packed" 2>&1 | tail -2
echo "== gemm new";  timeout 300 python tools/bench_kernels.py --only gemm 2>&1 | cut -c1-330
echo "== gemm base"; HRIEMO_LIB_PATH=tools/_build/libhriemo_base.so timeout 300 python tools/bench_kernels.py --only gemm 2>&1 | cut -c1-330
for i in 1 2; do
  echo "== new";  python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), 'gemm', round(d['roofline']['achieved']), 'attn', round(d['attention_roofline']['achieved']), d['clocks']['sm_mhz'])"
  echo "== base"; HRIEMO_LIB_PATH=tools/_build/libhriemo_base.so python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), 'gemm', round(d['roofline']['achieved']), 'attn', round(d['attention_roofline']['achieved']), d['clocks']['sm_mhz'])"
done
