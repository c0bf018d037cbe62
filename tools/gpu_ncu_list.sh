#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --batch 1024"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 234 -c 78 --csv --log-file gpurun_out/launches_v4.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"; tail -2 gpurun_out/ncu_launches.log
