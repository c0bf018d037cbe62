"""Micro-benchmark of the encoder attention backward (hriemo_attention_backward_bf16) per shape and implementation."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import ops
dev = "cuda"
impls = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "4"])]
for (B, H, Tq, Tk, dh) in [(512, 8, 500, 500, 96), (512, 8, 500, 64, 96), (512, 8, 64, 500, 96), (512, 8, 64, 64, 96), (512, 4, 300, 128, 64)]:
    d = H * dh
    q = torch.randn(B * Tq, d, device=dev).bfloat16(); k = torch.randn(B * Tk, d, device=dev).bfloat16(); v = torch.randn(B * Tk, d, device=dev).bfloat16()
    do = torch.randn(B * Tq, d, device=dev).bfloat16()
    out, lse = ops.attention(q, k, v, None, B, H, Tq, Tk, dh, want_lse=True)
    for impl in impls:
        fn = lambda: ops.attention_backward(q, k, v, out, do, lse, None, B, H, Tq, Tk, dh, impl=impl)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * H * Tq * Tk * dh * 7      # the two passes' seven tile GEMMs (S and dP are computed twice)
        print(dict(B=B, H=H, Tq=Tq, Tk=Tk, dh=dh, impl=impl, ms=round(ms, 4), tflops_7gemm=round(fl / ms / 1e9), tflops_5gemm=round(fl * 5 / 7 / ms / 1e9)), flush=True)
