"""One launch of each HBM-bound row kernel of the closing build at its north-star / training shape, bracketed by
cudaProfilerStart/Stop:  ncu --profile-from-start off --set full ... python tools/rows_once.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
B, T_a, L, d, H, dh = 2048, 500, 64, 768, 8, 96
a = torch.randn(B * T_a, d, device=dev, generator=g).bfloat16(); t = torch.randn(B * L, d, device=dev, generator=g).bfloat16()
vecs = [torch.rand(d, device=dev, generator=g) + 0.5 for _ in range(8)]
st = lambda x: torch.stack([x.float().mean(1), torch.rsqrt(x.float().var(1, unbiased=False) + 1e-5)], 1).contiguous()
sa, stt = st(a), st(t)
w = torch.sigmoid(torch.randn(B, d, device=dev, generator=g))
q = torch.randn(B * 4, d, device=dev).bfloat16(); kv = torch.randn(B * L, 2 * d, device=dev).bfloat16()
x = a[:256000]; dy = torch.randn(256000, d, device=dev).bfloat16()
qb = q[:512 * 4]; kvb = kv[:512 * L]; dob = torch.randn(512 * 4, d, device=dev).bfloat16()
def run():
    ops.ln_masked_mean(a, vecs[0], vecs[1], None, B, T_a, pre_ln=(vecs[4], vecs[5], sa))
    ops.gate_blend(a, T_a, t, (vecs[0], vecs[1]), (vecs[2], vecs[3]), w, B, L, pre_ln_a=(vecs[4], vecs[5], sa), pre_ln_t=(vecs[6], vecs[7], stt))
    ops.small_attention(q, kv[:, :d], kv[:, d:], None, B, H, 4, L, dh)
    ops.layernorm_backward(x, dy, vecs[0])
    ops.small_attention_backward(qb, kvb[:, :d], kvb[:, d:], dob, None, 512, H, 4, L, dh)
for _ in range(2): run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
