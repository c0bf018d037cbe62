#!/bin/bash
mkdir -p gpurun_out
python tools/sanitize_small.py > gpurun_out/sanitize_plain_v24.log 2>&1; echo "plain exit=$?"; tail -3 gpurun_out/sanitize_plain_v24.log
timeout 240 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitize_memcheck_v24.log 2>&1
echo "memcheck exit=$?"; grep -E "ERROR SUMMARY|all small-shape|FAILED|Error|error" gpurun_out/sanitize_memcheck_v24.log | head -5; tail -3 gpurun_out/sanitize_memcheck_v24.log | cut -c1-200
