#!/bin/bash
# ncu --set full on two GEMM launches of one forward at B=1024: the out-projection (+residual LN, stats)
# and the folded FFN first half (8th and 9th... pair-GEMM launches of the 4th forward).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --batch 1024"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s ${1:-120} -c ${2:-2} -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_gemm.log | cut -c1-200
