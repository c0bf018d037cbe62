#!/bin/bash
# Round-end check: every GPU test, smoke(), the contract bench + reference arm + ncu launch list (tools/gpu_full.sh)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu_full.sh
