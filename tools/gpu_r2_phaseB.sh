#!/bin/bash
# Round-2 phase B: the tf32-class mode's tests, the whole GPU suite, the default bench line (tf32x3 keys), the
# --precision tf32x3 line, the ncu launch list of one north-star forward and a full capture of the attention launches.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_precise_gpu.py -m gpu -x -q > gpurun_out/r2b_precise_tests.txt 2>&1; echo "precise exit=$?"; tail -15 gpurun_out/r2b_precise_tests.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gpu_tests.txt 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r2b_gpu_tests.txt
timeout 900 python bench.py --no-train > gpurun_out/r2b_bench_default.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/r2b_bench_default.log | cut -c1-3000
timeout 600 python bench.py --precision tf32x3 --steps 2 --no-torch --cpu-sample 256 > gpurun_out/r2b_bench_tf32x3.log 2>&1; echo "tf32x3 exit=$?"; tail -1 gpurun_out/r2b_bench_tf32x3.log | cut -c1-2500
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg --clock-control none -s 480 -c 156 --csv --log-file gpurun_out/r2b_launches.csv $CMD > gpurun_out/r2b_ncu_launches.log 2>&1
echo "launch list exit=$?"; tail -2 gpurun_out/r2b_ncu_launches.log | cut -c1-300
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train --batch 1024"
ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 24 -c 4 -f -o gpurun_out/r2b_prof_attn $CMD2 > gpurun_out/r2b_ncu_attn.log 2>&1
echo "ncu attn exit=$?"; tail -2 gpurun_out/r2b_ncu_attn.log | cut -c1-200
ls -la gpurun_out | grep r2b
