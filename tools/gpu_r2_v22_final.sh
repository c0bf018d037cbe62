#!/bin/bash
# Closing records of v22 on one box: every GPU test, smoke(), the default bench line, the reference arm, the ncu launch
# list of one eager training step.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/v22f_gpu_tests.txt 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/v22f_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/v22f_bench_default.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/v22f_bench_default.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v22f_bench_ref.log 2>&1; echo "ref exit=$?"; tail -1 gpurun_out/v22f_bench_ref.log | cut -c1-300
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg --clock-control none --csv --log-file gpurun_out/v22_train_launches.csv python tools/train_launch_list.py 512 > gpurun_out/v22_ncu_train.log 2>&1
echo "train launch list exit=$?"
