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
torch.cuda.synchronize()
print("FAILED: " + ", ".join(bad) if bad else "all small-shape checks passed")
sys.exit(1 if bad else 0)
