#!/bin/bash
# A / B of the GEMM tile walk (strided = default, contiguous = M-major ranges per owner): GEMM tests, the forward bench twice
# each, and the DRAM traffic of the GEMM launches of one forward under each walk.
mkdir -p gpurun_out
HRIEMO_GEMM_WALK=contiguous timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -x -q -k "gemm or golden" 2>&1 | tail -2
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train"
for i in 1 2; do
  $CMD 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('strided   ', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"
  HRIEMO_GEMM_WALK=contiguous $CMD 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('contiguous', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train"
HRIEMO_GEMM_WALK=contiguous ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg --clock-control none -s 450 -c 170 --csv --log-file gpurun_out/v24_launches_contiguous.csv $CMD > gpurun_out/v24_ncu_contig.log 2>&1
echo "launch list exit=$?"
