#!/bin/bash
# Round-2 phase C: decoder / gate kernels (decoder attention staged in shared memory, cp.async ring in the gate pooling,
# gate MLP layer 1 on the split-operand GEMM, precomputed layer-0 self-attention block): tests, bench, launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_precise_gpu.py -m gpu -x -q > gpurun_out/r2c_tests.txt 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/r2c_tests.txt
timeout 300 python tools/bench_kernels.py --only small 2>&1 | cut -c1-220
timeout 600 python bench.py --no-train --no-e2e --no-torch --cpu-sample 256 > gpurun_out/r2c_bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/r2c_bench.log | cut -c1-400
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg --clock-control none -s ${1:-450} -c 170 --csv --log-file gpurun_out/r2c_launches.csv $CMD > gpurun_out/r2c_ncu_launches.log 2>&1
echo "launch list exit=$?"
