"""Attention forward over a range of sequence lengths (one process per kernel selection: the HRIEMO_ATTN_FWD* variables are
read once): python tools/attn_sweep.py [dh H]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from hriemo import ops
from bench_kernels import timeit
dh = int(sys.argv[1]) if len(sys.argv) > 1 else 96
H = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = H * dh
for T in (160, 200, 256, 300, 320, 384, 448, 500, 576, 640, 768):
    B = max(32, 512 * 500 * 500 // (T * T) // 32 * 32)
    q = torch.randn(B * T, d, device="cuda").bfloat16(); k = torch.randn(B * T, d, device="cuda").bfloat16(); v = torch.randn(B * T, d, device="cuda").bfloat16()
    med, best = timeit(lambda: ops.attention(q, k, v, None, B, H, T, T, dh), iters=15)
    g = torch.Generator().manual_seed(T)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    pad = (torch.arange(T)[None, :] >= lens[:, None]).cuda()
    medm, _ = timeit(lambda: ops.attention(q, k, v, pad, B, H, T, T, dh), iters=15)
    print(f"T={T:5d} B={B:5d} dh={dh}: {med:.4f} ms ({4.0 * B * H * T * T * dh / med / 1e9:.0f} TFLOP/s), ragged mask {medm:.4f} ms", flush=True)
