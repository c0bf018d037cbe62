#!/bin/bash
# ncu launch list of one north-star forward at the final state of round 2
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train"
$CMD > gpurun_out/v25_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg --clock-control none -s 450 -c 176 --csv --log-file gpurun_out/v25_launches.csv $CMD > gpurun_out/v25_ncu_launches.log 2>&1
echo "launch list exit=$?"
