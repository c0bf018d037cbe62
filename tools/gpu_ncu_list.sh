#!/bin/bash
# ncu launch list of one forward at B=1024 (78 launches after 3 warm-up forwards): per-launch duration, DRAM bytes, tensor-pipe %.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --batch 1024"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg --clock-control none -s 234 -c 78 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"; tail -2 gpurun_out/ncu_launches.log | cut -c1-300
