#!/bin/bash
# A / B of the slab row budget of the forward (1 Mi rows = two slabs at the north star, 2 Mi + = one slab)
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train"
for i in 1 2; do
  $CMD 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1Mi rows ', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"
  HRIEMO_MAX_ROWS_PER_SLAB=2200000 $CMD 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('2.2M rows', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done
HRIEMO_MAX_ROWS_PER_SLAB=2200000 timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "full_size or golden" 2>&1 | tail -2
HRIEMO_MAX_ROWS_PER_SLAB=2200000 timeout 600 python bench.py --no-e2e --no-torch --no-ragged --no-train --cpu-sample 256 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('parity', d['parity']['logits_max_abs'], d['parity']['thr_agree_all'], d['value'])"
