#!/bin/bash
# v22 records of the other BASELINE configurations on one GPU: config 3 (MOSEI, the reference's published configuration)
# with the former gate / decoder kernels and with the closing ones, config 2 (IEMOCAP 300 / 50).
mkdir -p gpurun_out
C3="python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train"
HRIEMO_GATE_BLEND_V1=1 HRIEMO_DECODER_ATTN_V1=1 $C3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 former kernels', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
$C3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 v22           ', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
timeout 600 python bench.py --workload cfg3 --no-train > gpurun_out/v22_bench_cfg3_n1.log 2>&1; echo "cfg3 exit=$?"; tail -1 gpurun_out/v22_bench_cfg3_n1.log | cut -c1-200
timeout 600 python bench.py --workload cfg2 --no-train > gpurun_out/v22_bench_cfg2_n1.log 2>&1; echo "cfg2 exit=$?"; tail -1 gpurun_out/v22_bench_cfg2_n1.log | cut -c1-200
