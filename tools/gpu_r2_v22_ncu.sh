#!/bin/bash
# ncu --set full of the closing kernels: the new HBM-bound row kernels (one launch each, through the kernel bench) and the
# dominant GEMM (three launches of one forward at B = 1024: QKV with folded LayerNorm, out-projection + residual, FFN1).
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"decoder_attention_mma|gate_blend_stream|layernorm_backward_ring|ln_masked_mean|decoder_attention_backward" -f -o gpurun_out/v22_prof_rows python tools/rows_once.py > gpurun_out/v22_ncu_rows.log 2>&1
echo "ncu rows exit=$?"; tail -2 gpurun_out/v22_ncu_rows.log | cut -c1-200
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-torch --no-ragged --no-train --batch 1024"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s ${1:-120} -c 4 -f -o gpurun_out/v22_prof_gemm $CMD > gpurun_out/v22_ncu_gemm.log 2>&1
echo "ncu gemm exit=$?"; tail -2 gpurun_out/v22_ncu_gemm.log | cut -c1-200
