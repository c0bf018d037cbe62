#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python tools/bench_kernels.py --only small 2>&1 | cut -c1-330
for i in 1 2; do
  echo "== new";  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), 'gemm', round(d['roofline']['achieved']), 'attn', round(d['attention_roofline']['achieved']), d['clocks']['sm_mhz'])"
  echo "== base"; HRIEMO_LIB_PATH=tools/_build/libhriemo_base.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), 'gemm', round(d['roofline']['achieved']), 'attn', round(d['attention_roofline']['achieved']), d['clocks']['sm_mhz'])"
done
