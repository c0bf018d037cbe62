#!/bin/bash
# Same-box A/B of the forward bench, alternating in-tree library and tools/_build/libhriemo_base.so, $1 rounds of 8 steps
for i in $(seq 1 ${1:-3}); do
  for v in new base; do
    if [ $v = base ]; then export HRIEMO_LIB_PATH=tools/_build/libhriemo_base.so; else unset HRIEMO_LIB_PATH; fi
    python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['ms_per_step'],2), 'gemm', round(d['roofline']['achieved']), 'share', round(d['roofline']['share_of_step'],3), 'attn', round(d['attention_roofline']['achieved']), 'share', round(d['attention_roofline']['share_of_step'],3), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))"
  done
done
