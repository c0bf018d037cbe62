"""CPU oracle for the HRI-EMO fusion-and-decode forward path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`hri-emo_b200/`) may
import this file; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do, and only as the
checker / the CPU arm, never as the thing shipped.

What it is
----------
A from-scratch functional restatement of the forward arithmetic that the
reference composes out of `torch.nn` modules.  The arithmetic itself lives in a
third-party dependency that is NOT vendored under /root/reference: PyTorch
(`nn.MultiheadAttention`, `nn.LayerNorm`, `nn.Linear`, `torch.sigmoid`; the
reference does not pin it — README badge "2.0+", authors ran 2.9.0+cu126, this
image has 2.11.0+cu128).  The published algorithm restated here is
"Attention is all you need" multi-head attention with PyTorch's packed
in-projection layout ([Wq;Wk;Wv] rows, heads = contiguous column slices),
post-LN residual blocks, LayerNorm with biased variance and eps=1e-5.

torch is used here purely as an array library (matmul / exp / sum on CPU
tensors, float64 for checking, float32 for the timed CPU baseline); no
`torch.nn` module and no `F.multi_head_attention_forward` is called.

Parity pin
----------
The reference's own tests pin nothing but shapes (tests/*.py only print), so the
oracle is pinned against outputs of the reference itself, generated in the
authoring container by importing /root/reference (script:
tests/golden/make_golden.py, fixtures: tests/golden/*.pt) — see
tests/test_oracle_golden.py.  In float64 the oracle matches the imported
reference to ~1e-15.

Each function cites the reference file:line it follows (paths relative to the
reference repository root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor
State = Dict[str, Tensor]

LN_EPS = 1e-5  # nn.LayerNorm default, used by every LayerNorm in models/*.py


# --------------------------------------------------------------------------- #
# primitives (PyTorch semantics restated)
# --------------------------------------------------------------------------- #
def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """y = x W^T + b  (nn.Linear; weight is [out, in])."""
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = LN_EPS) -> Tensor:
    """nn.LayerNorm over the last dim: biased variance, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc / torch.sqrt(var + eps) * w + b


def relu(x: Tensor) -> Tensor:
    return torch.clamp(x, min=0)


def sigmoid(x: Tensor) -> Tensor:
    return 1.0 / (1.0 + torch.exp(-x))


def softmax_lastdim(s: Tensor) -> Tensor:
    """Numerically-stable softmax; a row that is all -inf yields NaN, like
    torch.softmax (the reference propagates those NaNs)."""
    m = s.max(dim=-1, keepdim=True).values
    e = torch.exp(s - m)
    return e / e.sum(dim=-1, keepdim=True)


def mha(
    sd: State,
    prefix: str,
    x_q: Tensor,
    x_kv: Tensor,
    n_heads: int,
    key_padding_mask: Optional[Tensor],
    need_weights: bool = False,
) -> Tuple[Tensor, Optional[Tensor]]:
    """nn.MultiheadAttention(batch_first=True) forward in eval mode.

    Call sites restated: models/cross_modal_block_tacfn.py:74-80, 85-91, 98-104,
    111-117; models/cross_modal_block.py:56-59, 64-67;
    models/emotion_decoder.py:42, 48-54.
    Packed in-projection rows [0:d]=Wq, [d:2d]=Wk, [2d:3d]=Wv; head h owns
    columns h*dh:(h+1)*dh; scores scaled by 1/sqrt(dh); PAD keys (True) get
    -inf; attention weights (if requested) are averaged over heads.
    """
    w = sd[prefix + "in_proj_weight"]
    b = sd[prefix + "in_proj_bias"]
    d = w.shape[1]
    dh = d // n_heads
    B, Tq, _ = x_q.shape
    Tk = x_kv.shape[1]
    q = linear(x_q, w[0:d], b[0:d])
    k = linear(x_kv, w[d : 2 * d], b[d : 2 * d])
    v = linear(x_kv, w[2 * d : 3 * d], b[2 * d : 3 * d])
    q = q.reshape(B, Tq, n_heads, dh).permute(0, 2, 1, 3)
    k = k.reshape(B, Tk, n_heads, dh).permute(0, 2, 1, 3)
    v = v.reshape(B, Tk, n_heads, dh).permute(0, 2, 1, 3)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)  # [B,H,Tq,Tk]
    if key_padding_mask is not None:
        neg = torch.full((), float("-inf"), dtype=s.dtype)
        s = torch.where(key_padding_mask[:, None, None, :], neg, s)
    p = softmax_lastdim(s)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B, Tq, d)
    out = linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])
    return out, (p.mean(dim=1) if need_weights else None)


def _ln(sd: State, prefix: str, x: Tensor) -> Tensor:
    return layer_norm(x, sd[prefix + "weight"], sd[prefix + "bias"])


def _ffn(sd: State, prefix: str, x: Tensor) -> Tensor:
    """nn.Sequential(Linear, ReLU, Linear): cross_modal_block_tacfn.py:43-52."""
    h = relu(linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"]))
    return linear(h, sd[prefix + "2.weight"], sd[prefix + "2.bias"])


# --------------------------------------------------------------------------- #
# encoder
# --------------------------------------------------------------------------- #
def cross_modal_block_tacfn(
    sd: State, p: str, a: Tensor, t: Tensor, m_a, m_t, n_heads: int, want_attn: bool = False
):
    """models/cross_modal_block_tacfn.py:62-125 (CrossModalBlock.forward), eval
    mode (dropout = identity).  Both cross-attentions read the post-self-
    attention streams."""
    maps = {}
    sa, w_ = mha(sd, p + "self_attn_a.", a, a, n_heads, m_a, want_attn)  # :74-80
    a_s = _ln(sd, p + "self_norm_a.", a + sa)  # :81
    maps["audio_self"] = w_
    st, w_ = mha(sd, p + "self_attn_t.", t, t, n_heads, m_t, want_attn)  # :85-91
    t_s = _ln(sd, p + "self_norm_t.", t + st)  # :92
    maps["text_self"] = w_
    a2t, w_ = mha(sd, p + "attn_a2t.", a_s, t_s, n_heads, m_t, want_attn)  # :98-104
    a1 = _ln(sd, p + "norm_a1.", a_s + a2t)  # :105
    a_o = _ln(sd, p + "norm_a2.", a1 + _ffn(sd, p + "ffn_a.", a1))  # :106
    maps["audio_queries_text"] = w_
    t2a, w_ = mha(sd, p + "attn_t2a.", t_s, a_s, n_heads, m_a, want_attn)  # :111-117
    t1 = _ln(sd, p + "norm_t1.", t_s + t2a)  # :118
    t_o = _ln(sd, p + "norm_t2.", t1 + _ffn(sd, p + "ffn_t.", t1))  # :119
    maps["text_queries_audio"] = w_
    return a_o, t_o, (maps if want_attn else None)


def cross_modal_block_legacy(sd: State, p: str, a: Tensor, t: Tensor, m_a, m_t, n_heads: int):
    """models/cross_modal_block.py:44-71: no intra-modal stage; both directions
    read the layer inputs."""
    a2t, _ = mha(sd, p + "attn_a2t.", a, t, n_heads, m_t)  # :56-59
    a1 = _ln(sd, p + "norm_a1.", a + a2t)  # :60
    a_o = _ln(sd, p + "norm_a2.", a1 + _ffn(sd, p + "ffn_a.", a1))  # :61
    t2a, _ = mha(sd, p + "attn_t2a.", t, a, n_heads, m_a)  # :64-67
    t1 = _ln(sd, p + "norm_t1.", t + t2a)  # :68
    t_o = _ln(sd, p + "norm_t2.", t1 + _ffn(sd, p + "ffn_t.", t1))  # :69
    return a_o, t_o


def _num_layers(sd: State, prefix: str) -> int:
    n = 0
    while any(k.startswith(f"{prefix}{n}.") for k in sd):
        n += 1
    return n


def cross_modal_transformer(
    sd: State, p: str, a, t, m_a, m_t, n_heads: int, want_attn: bool = False, legacy: bool = False
):
    """models/cross_modal_block_tacfn.py:146-166 / models/cross_modal_block.py:88-95."""
    attn: List[dict] = []
    for i in range(_num_layers(sd, p + "layers.")):
        lp = f"{p}layers.{i}."
        if legacy:
            a, t = cross_modal_block_legacy(sd, lp, a, t, m_a, m_t, n_heads)
        else:
            a, t, mp = cross_modal_block_tacfn(sd, lp, a, t, m_a, m_t, n_heads, want_attn)
            if want_attn:
                attn.append(mp)
    return a, t, (attn if want_attn else None)


# --------------------------------------------------------------------------- #
# gates
# --------------------------------------------------------------------------- #
def masked_mean(x: Tensor, mask: Optional[Tensor]) -> Tensor:
    """models/beta_gate_tacfn.py:6-24 (identical to models/beta_gate.py:6-33)."""
    if mask is None:
        return x.mean(dim=1)
    valid = (~mask).to(x.dtype)
    denom = torch.clamp(valid.sum(dim=1, keepdim=True), min=1.0)
    return (x * valid.unsqueeze(-1)).sum(dim=1) / denom


def _gate_input(a_pool: Tensor, t_pool: Tensor) -> Tensor:
    return torch.cat([a_pool, t_pool, (a_pool - t_pool).abs(), a_pool * t_pool], dim=-1)


def _gate_mlp(sd: State, p: str, g: Tensor) -> Tensor:
    h = relu(linear(g, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"]))
    return linear(h, sd[p + "mlp.2.weight"], sd[p + "mlp.2.bias"])


def beta_gate_tacfn(sd: State, p: str, a: Tensor, t: Tensor, m_a, m_t):
    """models/beta_gate_tacfn.py:68-118 (vector gate).  Fusion length = T_t;
    all T_a rows feed the pooled mean."""
    a_n = _ln(sd, p + "norm_a.", a)  # :79
    t_n = _ln(sd, p + "norm_t.", t)  # :80
    g = _gate_input(masked_mean(a_n, m_a), masked_mean(t_n, m_t))  # :83-89
    w = sigmoid(_gate_mlp(sd, p, g))  # :92   [B,d]
    beta = w.mean(dim=-1, keepdim=True)  # :95   [B,1]
    L = t_n.shape[1]  # :98-104
    a_n = a_n[:, :L]  # :107-110
    h = w[:, None, :] * a_n + (1.0 - w[:, None, :]) * t_n  # :113-116
    return h, beta


def beta_gate_legacy(sd: State, p: str, a: Tensor, t: Tensor, m_a, m_t):
    """models/beta_gate.py:60-114 (scalar gate, no LayerNorm)."""
    g = _gate_input(masked_mean(a, m_a), masked_mean(t, m_t))  # :81-87
    beta = sigmoid(_gate_mlp(sd, p, g))  # :90   [B,1]
    L = t.shape[1]  # :97-101
    h = beta[:, :, None] * a[:, :L] + (1.0 - beta[:, :, None]) * t[:, :L]  # :103-112
    return h, beta


def build_fused_mask(m_a, m_t, L: int):
    """models/fusion_with_emotion_decoder.py:71-115."""
    if m_a is None and m_t is None:
        return None

    def fit(m):
        if m is None:
            return None
        if m.shape[1] < L:
            pad = torch.ones(m.shape[0], L - m.shape[1], dtype=torch.bool)
            return torch.cat([m, pad], dim=1)
        return m[:, :L]

    ma, mt = fit(m_a), fit(m_t)
    if ma is None:
        return mt
    if mt is None:
        return ma
    return ma | mt


# --------------------------------------------------------------------------- #
# decoder
# --------------------------------------------------------------------------- #
def decoder_layer(sd: State, p: str, z: Tensor, mem: Tensor, mem_mask, n_heads: int, want_attn=False):
    """models/emotion_decoder.py:33-64 (ExplainableDecoderLayer.forward)."""
    s, _ = mha(sd, p + "self_attn.", z, z, n_heads, None)  # :42
    z = _ln(sd, p + "norm1.", z + s)  # :43
    c, w_ = mha(sd, p + "cross_attn.", z, mem, n_heads, mem_mask, want_attn)  # :48-54
    z = _ln(sd, p + "norm2.", z + c)  # :55
    f = linear(relu(linear(z, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
               sd[p + "linear2.weight"], sd[p + "linear2.bias"])  # :58
    z = _ln(sd, p + "norm3.", z + f)  # :59
    return z, w_


def emotion_decoder(sd: State, p: str, mem: Tensor, mem_mask, n_heads: int, want_attn=False):
    """models/emotion_decoder.py:117-162."""
    B = mem.shape[0]
    q = sd[p + "emotion_queries"].to(mem.dtype)
    z = q.unsqueeze(0).expand(B, -1, -1)  # :127
    attn = []
    for i in range(_num_layers(sd, p + "layers.")):
        z, w_ = decoder_layer(sd, f"{p}layers.{i}.", z, mem, mem_mask, n_heads, want_attn)
        if want_attn:
            attn.append(w_)
    logits = None
    if (p + "out_proj.weight") in sd:  # :153-155
        logits = linear(z, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"]).squeeze(-1)
    return z, logits, (attn if want_attn else None)


# --------------------------------------------------------------------------- #
# compositions
# --------------------------------------------------------------------------- #
def _ensure_3d(x: Tensor) -> Tensor:
    """models/fusion_with_emotion_decoder.py:60-69."""
    if x.dim() == 2:
        return x.unsqueeze(1)
    if x.dim() == 3:
        return x
    raise ValueError(f"Expected 2D or 3D tensor, got {x.shape}")


def cast_state(sd: State, dtype) -> State:
    return {k: (v.detach().to(dtype) if v.is_floating_point() else v.detach()) for k, v in sd.items()}


def fusion_with_emotion_decoder(
    sd: State, h_a: Tensor, h_t: Tensor, m_a=None, m_t=None, n_heads: int = 8,
    return_attention: bool = False, prefix: str = "", intermediates: Optional[dict] = None,
):
    """models/fusion_with_emotion_decoder.py:120-197."""
    a = _ensure_3d(h_a)
    t = _ensure_3d(h_t)
    a, t, enc_attn = cross_modal_transformer(sd, prefix + "cross_modal.", a, t, m_a, m_t, n_heads, return_attention)
    h, beta = beta_gate_tacfn(sd, prefix + "beta_gate.", a, t, m_a, m_t)  # :159
    fm = build_fused_mask(m_a, m_t, h.shape[1])  # :165
    z, logits, dec_attn = emotion_decoder(sd, prefix + "emotion_decoder.", h, fm, n_heads, return_attention)
    if intermediates is not None:
        intermediates.update(h_a_tilde=a, h_t_tilde=t, h_fusion=h, fused_mask=fm)
    if return_attention:
        return logits, beta, z, {"encoder": enc_attn, "decoder": dec_attn}
    return logits, beta, z


def mosei_fusion_with_emotion_decoder(
    sd: State, h_a: Tensor, h_t: Tensor, m_a=None, m_t=None, n_heads: int = 4,
    return_attention: bool = False,
):
    """models/mosei_fusion_with_emotion_decoder.py:55-79."""
    a = linear(h_a, sd["audio_proj.weight"], sd["audio_proj.bias"])  # :64
    t = linear(h_t, sd["text_proj.weight"], sd["text_proj.bias"])  # :66
    return fusion_with_emotion_decoder(sd, a, t, m_a, m_t, n_heads, return_attention, prefix="backbone.")


def fusion_classifier(sd: State, h_a: Tensor, h_t: Tensor, m_a=None, m_t=None, n_heads: int = 8):
    """models/fusion_classifier.py:98-150: encoder + vector gate + UNMASKED mean
    over time + LN/Linear/ReLU/Linear head."""
    a = _ensure_3d(h_a)
    t = _ensure_3d(h_t)
    a, t, _ = cross_modal_transformer(sd, "cross_modal.", a, t, m_a, m_t, n_heads)
    h, beta = beta_gate_tacfn(sd, "beta_gate.", a, t, m_a, m_t)
    pooled = h.mean(dim=1)  # :145
    x = _ln(sd, "classifier.0.", pooled)
    x = relu(linear(x, sd["classifier.1.weight"], sd["classifier.1.bias"]))
    logits = linear(x, sd["classifier.4.weight"], sd["classifier.4.bias"])
    return logits, beta, pooled


def legacy_block_and_gate(sd_block: State, sd_gate: State, h_a: Tensor, h_t: Tensor, m_a=None, m_t=None,
                          n_heads: int = 8):
    """tests/test_beta_gate.py:15-22: legacy CrossModalTransformer + scalar BetaGate."""
    a, t, _ = cross_modal_transformer(sd_block, "", h_a, h_t, m_a, m_t, n_heads, legacy=True)
    return beta_gate_legacy(sd_gate, "", a, t, m_a, m_t)


# --------------------------------------------------------------------------- #
# synthetic workload shared by tests and bench
# --------------------------------------------------------------------------- #
def ragged_masks(B: int, T: int, gen: torch.Generator) -> Tensor:
    """True = PAD; per-sample valid length in [T//2, T], at least one valid key."""
    lens = torch.randint(max(T // 2, 1), T + 1, (B,), generator=gen)
    return torch.arange(T)[None, :] >= lens[:, None]


def algorithmic_flops_per_utt(T_a, T_t, d, n_e, L_f, L_d, h_beta, ffn_dec=2048, d_a=None, d_t=None):
    """Closed-form matmul FLOPs (2*MAC) per utterance, SURVEY §8(d)."""
    L = T_t
    f = L_f * (32 * d * d * (T_a + T_t) + 4 * d * (T_a + T_t) ** 2)
    f += 10 * d * h_beta
    f += L_d * (12 * n_e * d * d + 4 * L * d * d + 4 * n_e * n_e * d + 4 * n_e * L * d + 4 * n_e * d * ffn_dec)
    f += 2 * n_e * d
    if d_a is not None:
        f += 2 * (T_a * d_a + T_t * d_t) * d
    return f
