#!/bin/bash
# attention op tests, then the same-box A/B against tools/_build/libhriemo_base.so and a short full bench A/B
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -3
bash tools/gpu_attn_ab.sh
for i in 1 2; do
  echo "== new";  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), 'gemm', round(d['roofline']['achieved']), 'attn', round(d['attention_roofline']['achieved']), d['clocks']['sm_mhz'])"
  echo "== base"; HRIEMO_LIB_PATH=tools/_build/libhriemo_base.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), 'gemm', round(d['roofline']['achieved']), 'attn', round(d['attention_roofline']['achieved']), d['clocks']['sm_mhz'])"
done
