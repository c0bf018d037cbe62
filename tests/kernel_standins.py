"""Torch stand-ins for the C-ABI kernels that hriemo/backward.py schedules — TEST INFRASTRUCTURE ONLY.

They let the CPU suite (`-m "not gpu"`) run the *schedule* of the backward pass (which kernel reads which saved
activation, which operand is transposed, where the residual branches join, which gradient lands under which
parameter name) against autograd without a GPU.  The kernels themselves are checked on the B200 by
tests/test_backward_gpu.py and tests/test_ops_gpu.py; nothing in the product package imports this file, and the
product has no CPU path: `install()` monkey-patches `hriemo.ops` / `hriemo.engine` inside a pytest fixture only.

Each stand-in keeps the kernel's contract (argument order, dtypes, which outputs are bf16) and computes in fp32."""
from __future__ import annotations

import math

import torch

bf16, f32 = torch.bfloat16, torch.float32
# install(..., exact=True) turns every dtype into float64 and every rounding into the identity: the schedule
# must then agree with autograd to round-off, which separates logic errors from bf16 noise.
LO, HI = bf16, f32
EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_RESID, EPI_BIAS_RESID_F32, _, EPI_BIAS_F32, EPI_BIAS_MASK = range(7)
ACT_NONE, ACT_RELU, ACT_SIGMOID = range(3)
EPS = 1e-5


def cast_bf16(x, ld_out=None):
    assert x.dtype == HI and x.dim() == 2
    return x.to(LO)


def gemm(a, w, bias, epilogue=EPI_BIAS, resid=None, out=None, cta_pair=0, a_ln=None, resid_ln=None, want_stats=False,
         tag="gemm"):
    assert a.dtype == LO and w.dtype == LO and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a_ln is None and resid_ln is None and not want_stats and out is None
    y = a.to(HI) @ w.to(HI).t()
    if bias is not None:
        y = y + bias
    if epilogue == EPI_BIAS_RELU:
        y = torch.relu(y)
    if epilogue in (EPI_BIAS_RESID, EPI_BIAS_RESID_F32):
        assert resid is not None and resid.dtype == (HI if epilogue == EPI_BIAS_RESID_F32 else LO)
        assert resid.shape == y.shape
        y = y + resid.to(HI)
    elif epilogue == EPI_BIAS_MASK:
        assert resid is not None and resid.dtype == LO and resid.shape == y.shape
        y = y * (resid > 0).to(HI)
    else:
        assert resid is None
    return y if epilogue in (EPI_BIAS_RESID_F32, EPI_BIAS_F32) else y.to(LO)


def split3(x, weight=False, relu=False):
    """hi + lo split operands of the tf32-class GEMM: the stand-in keeps the value in one piece (only the modules'
    prepared-operand caches call it on this path)."""
    return (torch.relu(x) if relu else x).to(LO)


def fold_ln_weight(w, bias, gamma, beta, *args, **kwargs):
    """Built by the modules' prepared-operand caches; the training schedule never reads the folded operands."""
    wf = (w * gamma[None, :]).to(LO)
    return wf, wf.to(HI).sum(1), (bias if bias is not None else 0) + w @ beta


def sgemm(a, w, bias, act=ACT_NONE):
    assert a.dtype == HI and w.dtype == HI
    if act & 4:
        a, act = torch.relu(a), act & 3
    y = a @ w.t() + (bias if bias is not None else 0)
    return torch.relu(y) if act == ACT_RELU else (torch.sigmoid(y) if act == ACT_SIGMOID else y)


def transpose_bf16(w):
    assert w.dtype == LO
    return w.t().contiguous()


def layernorm(x, gamma, beta, want_bf16=True, want_f32=False, eps=EPS):
    y = torch.nn.functional.layer_norm(x.to(HI), (x.shape[1],), gamma, beta, eps)
    return (y.to(LO) if want_bf16 else None), (y if want_f32 else None)


def _heads(x, B, T, H, dh):
    return x.to(HI).reshape(B, T, H, dh).transpose(1, 2)


def _drop_probs(p, drop, B, H, Tq, Tk):
    """Dropout on attention probabilities [B, H, Tq, Tk] with the masks of csrc/dropout.cuh (hriemo/dropout.py holds the
    same arithmetic): stream (b, h) has the key key_bh(key, b * H + h), rows are queries, columns keys."""
    if drop is None:
        return p
    from hriemo import dropout as D

    p8, scale, key = drop
    keep = torch.stack([D.keep_mask(Tq, Tk, D.key_bh(key, bh), p8) for bh in range(B * H)]).view(B, H, Tq, Tk)
    return p * keep.to(p.dtype) * scale


def dropout(x, drop, resid=None):
    from hriemo import dropout as D

    p8, scale, key = drop
    y = x.to(HI) * D.keep_mask(x.shape[0], x.shape[1], key, p8).to(HI) * scale
    if resid is not None:
        assert resid.dtype == x.dtype and resid.shape == x.shape
        y = y + resid.to(HI)
    return y.to(x.dtype)


def small_attention(q, k, v, key_pad, B, H, Nq, Tk, dh, want_probs=False, drop=None):
    s = _heads(q, B, Nq, H, dh) @ _heads(k, B, Tk, H, dh).transpose(-1, -2) / math.sqrt(dh)
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (_drop_probs(p, drop, B, H, Nq, Tk) @ _heads(v, B, Tk, H, dh)).transpose(1, 2).reshape(B * Nq, H * dh)
    return o.to(LO), (p.mean(1) if want_probs else None)


def small_attention_backward(q, k, v, d_out, key_pad, B, H, Nq, Tk, dh, out=None, drop=None):
    qr, kr, vr = (t.to(HI).clone().requires_grad_(True) for t in (q, k, v))
    s = _heads(qr, B, Nq, H, dh) @ _heads(kr, B, Tk, H, dh).transpose(-1, -2) / math.sqrt(dh)
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    o = (_drop_probs(torch.softmax(s, dim=-1), drop, B, H, Nq, Tk) @ _heads(vr, B, Tk, H, dh)).transpose(1, 2).reshape(B * Nq, H * dh)
    o.backward(d_out.to(HI))
    res = (qr.grad.to(LO), kr.grad.to(LO), vr.grad.to(LO))
    if out is None:
        return res
    for dst, src in zip(out, res):
        assert dst.dtype == LO and dst.shape == src.shape
        dst.copy_(src)
    return out


def _attn(qr, kr, vr, key_pad, B, H, Tq, Tk, dh, drop=None):
    s = _heads(qr, B, Tq, H, dh) @ _heads(kr, B, Tk, H, dh).transpose(-1, -2) / math.sqrt(dh)
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    p = _drop_probs(torch.softmax(s, dim=-1), drop, B, H, Tq, Tk)
    return (p @ _heads(vr, B, Tk, H, dh)).transpose(1, 2).reshape(B * Tq, H * dh), s


def attention(q, k, v, key_pad, B, H, Tq, Tk, dh, skip_padded_tiles=True, pair_heads=True, want_lse=False, out=None, drop=None):
    o, s = _attn(q, k, v, key_pad, B, H, Tq, Tk, dh, drop)
    return (o.to(LO), torch.logsumexp(s, dim=-1)) if want_lse else o.to(LO)


def attention_backward(q, k, v, out, d_out, lse, key_pad, B, H, Tq, Tk, dh, grads=None, impl=0, drop=None):
    assert lse.shape == (B, H, Tq) and out.shape == d_out.shape == (B * Tq, H * dh)
    qr, kr, vr = (t.to(HI).clone().requires_grad_(True) for t in (q, k, v))
    o, _ = _attn(qr, kr, vr, key_pad, B, H, Tq, Tk, dh, drop)
    o.backward(d_out.to(HI))
    res = (qr.grad.to(LO), kr.grad.to(LO), vr.grad.to(LO))
    if grads is None:
        return res
    for dst, src in zip(grads, res):
        assert dst.dtype == LO and dst.shape == src.shape
        dst.copy_(src)
    return grads


def ln_masked_mean(x, gamma, beta, pad, B, T, apply_ln=True, eps=EPS, pre_ln=None):
    assert pre_ln is None
    y = x.to(HI)
    if apply_ln:
        y = torch.nn.functional.layer_norm(y, (x.shape[1],), gamma, beta, eps)
    y = y.view(B, T, -1)
    if pad is None:
        return y.mean(1)
    valid = (~pad).to(HI)
    return (y * valid[..., None]).sum(1) / valid.sum(1, keepdim=True).clamp(min=1.0)


def gate_input(a_pool, t_pool):
    return torch.cat([a_pool, t_pool, (a_pool - t_pool).abs(), a_pool * t_pool], dim=-1)


def gate_blend(a, T_a, t, ln_a, ln_t, w, B, L, apply_ln=True, w_is_scalar=False, want_bf16=True, want_f32=False,
               eps=EPS, pre_ln_a=None, pre_ln_t=None):
    assert not apply_ln and pre_ln_a is None and pre_ln_t is None and not w_is_scalar
    d = a.shape[1]
    h = w[:, None, :] * a.to(HI).view(B, T_a, d)[:, :L] + (1 - w[:, None, :]) * t.to(HI).view(B, L, d)
    h = h.reshape(B * L, d)
    return (h.to(LO) if want_bf16 else None), (h if want_f32 else None), w.mean(-1, keepdim=True)


def bce_beta_loss(logits, labels, beta, beta_weight=0.01, want_grads=True):
    x = logits.clone().requires_grad_(True)
    b = beta.clone().requires_grad_(True)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(x, labels) - beta_weight * (b * (1 - b)).mean()
    loss.backward()
    return loss.detach().view(1), x.grad, b.grad


def linear_wgrad(dy, x, dw=None, db=None, want_bias=True, accumulate=False):
    assert dy.dtype == LO and x.dtype == LO and dy.shape[0] == x.shape[0]
    assert dy.shape[1] % 128 == 0 and x.shape[1] % 128 == 0, "the tcgen05 wgrad kernel needs N, K multiples of 128"
    gw = dy.to(HI).t() @ x.to(HI)
    gb = dy.to(HI).sum(0)
    if dw is None:
        dw = torch.zeros_like(gw)
    assert dw.dtype == HI and dw.shape == gw.shape and dw.is_contiguous()
    dw.copy_(dw + gw if accumulate else gw)
    if not want_bias:
        return dw, None
    if db is None:
        db = torch.zeros_like(gb)
    assert db.shape == gb.shape and db.is_contiguous()
    db.copy_(db + gb if accumulate else gb)
    return dw, db


def linear_backward(dy, x, w_t, want_bias=True, relu_input=False):
    assert w_t.shape == (x.shape[1], dy.shape[1]), "w_t must be the transposed weight [K, N]"
    dx = gemm(dy, w_t, None, EPI_BIAS_MASK, resid=x) if relu_input else gemm(dy, w_t, None, EPI_BIAS)
    dw, db = linear_wgrad(dy, x, want_bias=want_bias)
    return dx, dw, db


def relu_backward(dy, h):
    assert dy.dtype == LO and h.dtype == LO and dy.shape == h.shape
    return torch.where(h > 0, dy, torch.zeros_like(dy))


def layernorm_backward(x, dy, gamma, dgamma=None, dbeta=None, accumulate=False, eps=EPS):
    assert x.dtype == LO and dy.dtype == LO and x.shape == dy.shape and not accumulate
    xr = x.to(HI).requires_grad_(True)
    g = gamma.clone().requires_grad_(True)
    b = torch.zeros_like(gamma).requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (x.shape[1],), g, b, eps).backward(dy.to(HI))
    return xr.grad.to(LO), g.grad, b.grad


def linear_backward_f32(dy, x, w, want_dx=True, want_dw=True, want_bias=True, dw=None, db=None, accumulate=False):
    assert dy.dtype == HI and not accumulate
    return (dy @ w if want_dx else None), (dy.t() @ x if want_dw else None), (dy.sum(0) if want_bias else None)


def act_backward_f32(dy, y, act):
    return dy * (y > 0).to(HI) if act == ACT_RELU else (dy * y * (1 - y) if act == ACT_SIGMOID else dy.clone())


def sum_rows(x, out=None, accumulate=False):
    assert out is None
    return x.to(HI).sum(0)


def gate_input_backward(dg, a_pool, t_pool):
    d = a_pool.shape[1]
    sg = torch.sign(a_pool - t_pool)
    g0, g1, g2, g3 = dg[:, :d], dg[:, d:2 * d], dg[:, 2 * d:3 * d], dg[:, 3 * d:]
    return g0 + sg * g2 + t_pool * g3, g1 - sg * g2 + a_pool * g3


def mask_inv_counts(like, pad, B, T):
    if pad is None:
        return torch.full((B,), 1.0 / T, dtype=HI)
    return 1.0 / (~pad).sum(1).clamp(min=1).to(HI)


def gate_blend_backward_w(dh, na, T_a, nt, dbeta, B, L):
    d = dh.shape[1]
    diff = na.to(HI).view(B, T_a, d)[:, :L] - nt.to(HI).view(B, L, d)
    dw = (dh.to(HI).view(B, L, d) * diff).sum(1)
    return dw + (dbeta.view(B, 1) / d if dbeta is not None else 0)


def gate_stream_grad(dh, L, w, one_minus, dpool, pad, inv_counts, B, T):
    d = dh.shape[1]
    coef = (1 - w) if one_minus else w
    dn = torch.zeros(B, T, d, dtype=HI)
    dn[:, :L] = coef[:, None, :] * dh.to(HI).view(B, L, d)
    valid = torch.ones(B, T, dtype=HI) if pad is None else (~pad).to(HI)
    dn = dn + valid[..., None] * (dpool * inv_counts[:, None])[:, None, :]
    return dn.view(B * T, d).to(LO)


def grad_norm_clip(grads, max_norm):
    total = grads.norm()
    return torch.stack([total, torch.clamp(max_norm / (total + 1e-6), max=1.0)])


def adamw_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
               grad_scale=None, params_bf16=None):
    g = grads * (grad_scale[0] if grad_scale is not None else 1.0)
    params.mul_(1 - lr * weight_decay)
    exp_avg.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    exp_avg_sq.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    denom = exp_avg_sq.sqrt() / math.sqrt(1 - betas[1] ** step) + eps
    params.addcdiv_(exp_avg, denom, value=-lr / (1 - betas[0] ** step))


def install(monkeypatch, exact: bool = False):
    """Replace the kernel wrappers the backward schedule uses with the stand-ins above (pytest monkeypatch).
    exact: all tensors float64, no rounding anywhere (parameters must be float64 too)."""
    import sys

    from hriemo import backward, engine, ops

    me = sys.modules[__name__]
    monkeypatch.setattr(me, "LO", torch.float64 if exact else bf16)
    monkeypatch.setattr(me, "HI", torch.float64 if exact else f32)
    for name, fn in list(vars(me).items()):
        if callable(fn) and not name.startswith("_") and name != "install" and hasattr(ops, name):
            monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(engine, "require_cuda", lambda x, what: None)
    if exact:
        monkeypatch.setattr(engine, "v32", lambda v: v.detach().double().contiguous())
        monkeypatch.setattr(engine, "w16", lambda w, k_pad=None: w.detach().double().contiguous())
        monkeypatch.setattr(backward, "bf16", torch.float64)
        monkeypatch.setattr(backward, "f32", torch.float64)
        from hriemo import train
        monkeypatch.setattr(train, "f32", torch.float64)
