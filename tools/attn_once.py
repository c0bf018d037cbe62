"""One attention launch per call shape (for ncu): python tools/attn_once.py B H Tq Tk dh [iters]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import ops
B, H, Tq, Tk, dh = (int(x) for x in sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
d = H * dh
q = torch.randn(B * Tq, d, device="cuda").bfloat16(); k = torch.randn(B * Tk, d, device="cuda").bfloat16(); v = torch.randn(B * Tk, d, device="cuda").bfloat16()
for _ in range(iters):
    ops.attention(q, k, v, None, B, H, Tq, Tk, dh)
torch.cuda.synchronize()
print("done")
