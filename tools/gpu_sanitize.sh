#!/bin/bash
# compute-sanitizer over the small-shape run of every pipeline kernel; summaries into gpurun_out/
mkdir -p gpurun_out
python tools/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain exit=$?"; tail -2 gpurun_out/sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all small-shape|FAILED" gpurun_out/sanitize_$tool.log | tail -3
done
