"""Small shapes of every pipeline kernel, once each, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
GEMM one-CTA and CTA-pair forms (incl. residual + LayerNorm epilogues), the four attention forward forms (general, DEFER,
paired heads, paired + DEFER) with and without masks, wgrad, the tcgen05 attention backward (both passes).  Results are
checked loosely against torch so that a silent corruption also fails the run."""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hri-emo_b200"))
from hriemo import lib as L, ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(3)
def rnd(*shape, dtype=torch.bfloat16): return (torch.randn(*shape, device=dev, generator=g)).to(dtype)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-9))
bad = []
def check(name, got, ref, tol=2e-2):
    r = rel(got, ref)
    print(f"{name:55s} rel err {r:.2e}")
    if not (r <= tol): bad.append(name)
# ---- GEMM
for pair in (1, 2):
    M, N, K = 300, 256, 128
    a, w, b = rnd(M, K), rnd(N, K), rnd(N, dtype=torch.float32)
    check(f"gemm bias cta_pair={pair}", ops.gemm(a, w, b, L.EPI_BIAS, cta_pair=pair), a.float() @ w.float().t() + b)
    r = rnd(M, N)
    out, stats = ops.gemm(a, w, b, L.EPI_BIAS_RESID, resid=r, want_stats=True, cta_pair=pair)
    ref = a.float() @ w.float().t() + b + r.float()
    check(f"gemm bias+resid+stats cta_pair={pair}", out, ref)
    check(f"gemm row means cta_pair={pair}", stats[:, 0], ref.mean(1), 5e-2)

# ---- attention forward, four forms
def attn_ref(q, k, v, pad, B, H, Tq, Tk, dh):
    qh = q.float().view(B, Tq, H, dh).transpose(1, 2); kh = k.float().view(B, Tk, H, dh).transpose(1, 2); vh = v.float().view(B, Tk, H, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    if pad is not None: s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B * Tq, H * dh)
for (B, H, Tq, Tk, dh, name) in [(2, 2, 300, 300, 96, "general"), (2, 2, 300, 64, 96, "DEFER"), (2, 2, 64, 300, 96, "paired"), (2, 2, 64, 64, 64, "paired+DEFER")]:
    d = H * dh
    q, k, v = rnd(B * Tq, d), rnd(B * Tk, d), rnd(B * Tk, d)
    for masked in (False, True):
        pad = None
        if masked:
            lens = torch.randint(Tk // 2, Tk + 1, (B, 1), device=dev, generator=g)
            pad = torch.arange(Tk, device=dev)[None, :] >= lens
        out, lse = ops.attention(q, k, v, pad, B, H, Tq, Tk, dh, want_lse=True)
        check(f"attention fwd {name} masked={masked}", out, attn_ref(q, k, v, pad, B, H, Tq, Tk, dh))
        if name in ("general", "paired"):
            do = rnd(B * Tq, d)
            dq, dk, dv = ops.attention_backward(q, k, v, out, do, lse, pad, B, H, Tq, Tk, dh)
            qr, kr, vr = (x.double().requires_grad_(True) for x in (q, k, v))
            qh = qr.view(B, Tq, H, dh).transpose(1, 2); kh = kr.view(B, Tk, H, dh).transpose(1, 2); vh = vr.view(B, Tk, H, dh).transpose(1, 2)
            s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
            if pad is not None: s = s.masked_fill(pad[:, None, None, :], float("-inf"))
            (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B * Tq, d).backward(do.double())
            check(f"attention bwd (tcgen05) {name} masked={masked} dq", dq, qr.grad)
            check(f"attention bwd (tcgen05) {name} masked={masked} dk", dk, kr.grad)
            check(f"attention bwd (tcgen05) {name} masked={masked} dv", dv, vr.grad)
# ---- wgrad
M, N, K = 700, 256, 128
dy, x = rnd(M, N), rnd(M, K)
res = ops.linear_wgrad(dy, x)
dw = res[0] if isinstance(res, (tuple, list)) else res
check("linear wgrad", dw, dy.float().t() @ x.float())
# ---- round-2 closing row kernels: warp-private decoder attention (+ its staged backward), streaming gate blend, gate pooling
#      with staged statistics / PAD bytes, LayerNorm backward ring, few-column fp32 GEMM
for (B, H, Nq, Tk, dh) in [(5, 8, 4, 50, 96), (3, 4, 6, 128, 64)]:
    d = H * dh
    q, kv = rnd(B * Nq, d), rnd(B * Tk, 2 * d)
    pad = torch.arange(Tk, device=dev)[None, :] >= torch.randint(Tk // 2, Tk + 1, (B, 1), device=dev, generator=g)
    out, _ = ops.small_attention(q, kv[:, :d], kv[:, d:], pad, B, H, Nq, Tk, dh)
    check(f"decoder attention {Nq}x{Tk}x{dh}", out, attn_ref(q, kv[:, :d], kv[:, d:], pad, B, H, Nq, Tk, dh))
    do = rnd(B * Nq, d)
    dq, dk, dv = ops.small_attention_backward(q, kv[:, :d], kv[:, d:], do, pad, B, H, Nq, Tk, dh)
    qr, kr, vr = (x.double().requires_grad_(True) for x in (q, kv[:, :d].contiguous(), kv[:, d:].contiguous()))
    sc = (qr.view(B, Nq, H, dh).transpose(1, 2) @ kr.view(B, Tk, H, dh).transpose(1, 2).transpose(-1, -2)) / math.sqrt(dh)
    sc = sc.masked_fill(pad[:, None, None, :], float("-inf"))
    (torch.softmax(sc, -1) @ vr.view(B, Tk, H, dh).transpose(1, 2)).transpose(1, 2).reshape(B * Nq, d).backward(do.double())
    check(f"decoder attention backward {Nq}x{Tk}x{dh} dq", dq, qr.grad)
    check(f"decoder attention backward {Nq}x{Tk}x{dh} dk", dk, kr.grad)
    check(f"decoder attention backward {Nq}x{Tk}x{dh} dv", dv, vr.grad)
F = torch.nn.functional
B, Ta, Lf, d = 5, 70, 20, 768
xa, xt = rnd(B * Ta, d), rnd(B * Lf, d)
vec = [torch.rand(d, device=dev, generator=g) + 0.5 for _ in range(8)]
stf = lambda x: torch.stack([x.float().mean(1), torch.rsqrt(x.float().var(1, unbiased=False) + 1e-5)], 1).contiguous()
w = torch.sigmoid(rnd(B, d, dtype=torch.float32))
ya = F.layer_norm(F.layer_norm(xa.float(), (d,), vec[4], vec[5], 1e-5), (d,), vec[0], vec[1], 1e-5).view(B, Ta, d)
yt = F.layer_norm(F.layer_norm(xt.float(), (d,), vec[6], vec[7], 1e-5), (d,), vec[2], vec[3], 1e-5).view(B, Lf, d)
_, hf, _ = ops.gate_blend(xa, Ta, xt, (vec[0], vec[1]), (vec[2], vec[3]), w, B, Lf, want_bf16=False, want_f32=True,
                          pre_ln_a=(vec[4], vec[5], stf(xa)), pre_ln_t=(vec[6], vec[7], stf(xt)))
check("gate blend (streaming)", hf.view(B, Lf, d), w[:, None] * ya[:, :Lf] + (1 - w[:, None]) * yt, 1e-4)
pad = torch.arange(Ta, device=dev)[None, :] >= torch.randint(Ta // 2, Ta + 1, (B, 1), device=dev, generator=g)
pooled = ops.ln_masked_mean(xa, vec[0], vec[1], pad, B, Ta, pre_ln=(vec[4], vec[5], stf(xa)))
valid = (~pad).float()[:, :, None]
check("gate pooling (staged statistics, masks)", pooled, (ya * valid).sum(1) / valid.sum(1).clamp(min=1), 1e-4)
x, dy = rnd(1001, d), rnd(1001, d)
dx, dg, db = ops.layernorm_backward(x, dy, vec[0])
xr = x.double().requires_grad_(True); gr = vec[0].double().requires_grad_(True); br = torch.zeros(d, device=dev, dtype=torch.float64, requires_grad=True)
F.layer_norm(xr, (d,), gr, br, 1e-5).backward(dy.double())
check("layernorm backward ring dx", dx, xr.grad)
check("layernorm backward ring dgamma", dg, gr.grad, 1e-3)
check("layernorm backward ring dbeta", db, br.grad, 1e-3)
a32, w32 = rnd(37, 768, dtype=torch.float32), rnd(1, 768, dtype=torch.float32)
check("sgemm few columns", ops.sgemm(a32, w32, None, L.ACT_NONE), a32 @ w32.t(), 1e-5)
torch.cuda.synchronize()
print("FAILED: " + ", ".join(bad) if bad else "all small-shape checks passed")
sys.exit(1 if bad else 0)
