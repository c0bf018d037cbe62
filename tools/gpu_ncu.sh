#!/bin/bash
# ncu evidence for profiles/: launch list of the bench step + full captures of the two tensor-core kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --batch 1024"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_fwd2 -s 16 -c 2 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn exit=$?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 76 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm exit=$?"
tail -3 gpurun_out/plain.log; ls -la gpurun_out/
