#!/bin/bash
# ncu --set full on the weight-gradient kernel (FFN1 shape of a 512-utterance slab)
mkdir -p gpurun_out
cat > /tmp/wg.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "hri-emo_b200"))
from hriemo import ops
dy = torch.randn(256000, 3072, device="cuda").bfloat16(); x = torch.randn(256000, 768, device="cuda").bfloat16()
for _ in range(3): ops.linear_wgrad(dy, x, want_bias=False)
torch.cuda.synchronize()
PY
python /tmp/wg.py && ncu --set full --clock-control none --import-source on -k regex:gemm_wgrad -s 2 -c 1 -f -o gpurun_out/prof_wgrad python /tmp/wg.py > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu exit=$?"
