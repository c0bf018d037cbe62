#!/bin/bash
# Round-2 phase A check: GPU test suite, the new default bench line (parity / gpu_eager_baseline / ragged / e2e stats /
# train_step), the stock-PyTorch arm and the CPU reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.txt 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/r2_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r2_bench_default.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/r2_bench_default.log | cut -c1-6000
timeout 600 python bench.py --impl torch_gpu --steps 3 > gpurun_out/r2_bench_torch_gpu.log 2>&1; echo "torch_gpu exit=$?"; tail -1 gpurun_out/r2_bench_torch_gpu.log | cut -c1-2500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.log 2>&1; echo "ref exit=$?"; tail -1 gpurun_out/r2_bench_ref.log | cut -c1-1500
