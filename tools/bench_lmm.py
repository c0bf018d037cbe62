import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hri-emo_b200"))
from hriemo import ops
B, T, d = 2048, 500, 768
x = torch.randn(B * T, d, device="cuda").bfloat16()
g, b = torch.rand(d, device="cuda") + 0.5, torch.randn(d, device="cuda")
g2, b2 = torch.rand(d, device="cuda") + 0.5, torch.randn(d, device="cuda")
st = torch.stack([x.float().mean(1), torch.rsqrt(x.float().var(1, unbiased=False) + 1e-5)], 1).contiguous()
pad = (torch.arange(T, device="cuda")[None, :] >= torch.randint(T // 2, T + 1, (B, 1), device="cuda"))
for name, kw in (("single LN", {}), ("pending LN + stats", dict(pre_ln=(g2, b2, st))), ("pending LN, no stats", dict(pre_ln=(g2, b2)))):
    for m in (None, pad):
        for _ in range(3): ops.ln_masked_mean(x, g, b, m, B, T, **kw)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): ops.ln_masked_mean(x, g, b, m, B, T, **kw)
        e.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(e) / 10
        frac = 1.0 if m is None else float((~pad).float().mean())
        print(f"{name:22s} mask={'yes' if m is not None else 'no ':3s} {ms:.3f} ms  {B*T*d*2*frac/ms/1e6:.0f} GB/s of rows read")
