#!/bin/bash
# v20: same-box A / B of the forward with the former gate / decoder kernels, ncu launch lists of one forward and of one training step.
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train"
for i in 1 2; do
  HRIEMO_GATE_BLEND_V1=1 HRIEMO_DECODER_ATTN_V1=1 $CMD 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('former kernels', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
  $CMD 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('v20           ', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg --clock-control none -s ${1:-450} -c 170 --csv --log-file gpurun_out/v20_launches.csv $CMD > gpurun_out/v20_ncu_launches.log 2>&1
echo "launch list exit=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg --clock-control none --csv --log-file gpurun_out/v20_train_launches.csv python tools/train_launch_list.py 512 > gpurun_out/v20_ncu_train.log 2>&1
echo "train launch list exit=$?"
