#!/bin/bash
# Full contract run: default bench line, reference arm, and the ncu launch list of one north-star forward.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-3000
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-1200
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg --clock-control none -s 468 -c 156 --csv --log-file gpurun_out/launches_ns.csv $CMD > gpurun_out/ncu_launches_ns.log 2>&1
echo "launch list exit=$?"
