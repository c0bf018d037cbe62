"""Kernel schedules of the fusion-and-decode forward path.

The nn.Modules in ``models/`` own the parameters (reference names and shapes);
this file holds (a) the prepared bf16 / concatenated operand sets derived from
them and (b) the order in which the C-ABI kernels are enqueued for each
reference sub-module.  Activations of the two encoder streams live in HBM as
row-major bf16 ``[B*T, d]`` matrices; the tiny decoder keeps an fp32 residual
stream next to a bf16 copy (GEMM operand).  See DESIGN.md.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import lib as L
from . import ops

bf16 = torch.bfloat16
f32 = torch.float32


# --------------------------------------------------------------------------- #
# activations
# --------------------------------------------------------------------------- #
@dataclass
class LazyLN:
    """A LayerNorm that has not been applied yet: the stream holds the pre-LayerNorm sum and every
    consumer applies (x - mean) * rstd * gamma + beta itself (fused GEMM epilogues, gate kernels)."""
    gamma: torch.Tensor   # fp32 [d]
    beta: torch.Tensor    # fp32 [d]
    stats: torch.Tensor   # fp32 [B*T, 2] = (mean, rstd) per row


@dataclass
class Seq:
    """A batch of sequences as one row-major matrix: x[b*T + t, :].  When `ln` is set, the logical
    value of the stream is LN(x) and x is the pre-LayerNorm tensor."""
    x: torch.Tensor                      # bf16 [B*T, d]
    B: int
    T: int
    x32: Optional[torch.Tensor] = None   # fp32 copy when the caller asked for one
    ln: Optional[LazyLN] = None

    @property
    def d(self) -> int:
        return self.x.shape[1]

    def as_f32_3d(self) -> torch.Tensor:
        if self.x32 is None:
            raise L.HriemoError("internal: fp32 output was not requested")
        return self.x32.view(self.B, self.T, -1)


def require_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise L.HriemoError(
            f"{what}: tensor is on {x.device}; the HRI-EMO B200 path has no CPU fallback "
            "(move the module and its inputs to a CUDA device)")


def to_seq(x: torch.Tensor, what: str, ld: Optional[int] = None) -> Seq:
    """[B,T,d] features (fp32 as the reference loads them, or bf16) -> bf16 Seq."""
    require_cuda(x, what)
    B, T, d = x.shape
    if x.dtype == bf16 and (ld is None or ld == d):
        return Seq(x.contiguous().view(B * T, d), B, T)
    if x.dtype != f32:
        x = x.float()
    x2 = x.contiguous().view(B * T, d)
    return Seq(ops.cast_bf16(x2, ld), B, T)


def check_mask(mask: Optional[torch.Tensor], B: int, T: int, name: str) -> Optional[torch.Tensor]:
    if mask is None:
        return None
    if mask.dim() != 2 or mask.shape[0] != B or mask.shape[1] != T:
        raise RuntimeError(f"{name}: key_padding_mask shape {tuple(mask.shape)} does not match ({B}, {T})")
    if mask.dtype != torch.bool:
        mask = mask != 0
    return mask.contiguous()


_warned_train = False


def warn_if_training(module: torch.nn.Module, p_drop: float) -> None:
    """Dropout lives in the TRAINING schedule (hriemo/backward.py, entered through the autograd boundary of the top-level
    models or hriemo.train.Trainer).  A forward in train() mode that does not go there -- gradients disabled, a sub-module
    called on its own, return_attention=True -- runs the inference schedule, which applies none: said once."""
    global _warned_train
    if module.training and p_drop > 0 and not _warned_train:
        _warned_train = True
        warnings.warn(
            "hri-emo_b200: this forward runs the inference schedule (no dropout, no autograd graph) although the module is "
            "in training mode; dropout is applied by the training path: model(...) of FusionWithEmotionDecoder / "
            "MoseiFusionWithEmotionDecoder with gradients enabled, or hriemo.train.Trainer.  Call .eval() for inference.",
            stacklevel=3)


# --------------------------------------------------------------------------- #
# prepared operands (bf16 weights, fused projections), cached per module
# --------------------------------------------------------------------------- #
class Prepared:
    """Caches tensors derived from a module's parameters; rebuilt when any parameter
    is replaced or modified in place (load_state_dict, optimizer step, .to())."""

    def __init__(self, module: torch.nn.Module, builder):
        self._module = module
        self._builder = builder
        self._key = None
        self._value = None

    def get(self):
        params = list(self._module.parameters())
        key = tuple((p.data_ptr(), p._version, p.device) for p in params)
        if key != self._key:
            for p in params:
                require_cuda(p, type(self._module).__name__ + " parameter")
            with torch.no_grad():
                self._value = self._builder()
            self._key = key
        return self._value


def w16(w: torch.Tensor, k_pad: Optional[int] = None) -> torch.Tensor:
    """nn.Linear weight [N,K] fp32 -> bf16 GEMM operand (K optionally zero-padded)."""
    w = w.detach()
    if w.dtype != f32:
        w = w.float()
    return ops.cast_bf16(w.contiguous(), k_pad)


def v32(v: torch.Tensor) -> torch.Tensor:
    v = v.detach()
    return v.contiguous() if v.dtype == f32 else v.float().contiguous()


def prep_folded(w: torch.Tensor, b: Optional[torch.Tensor], ln_module) -> dict:
    """Operands of a Linear fed by LN(x) when only the pre-LayerNorm x is in memory
    (hriemo_fold_ln_weight): w*gamma in bf16, its row sums, and b + w.beta."""
    wf, cs, bf = ops.fold_ln_weight(v32(w), v32(b) if b is not None else None, v32(ln_module.weight),
                                    v32(ln_module.bias))
    return dict(w=wf, b=bf, colsum=cs)


def cross_pair_weights(mha_q, mha_kv):
    d = mha_q.in_proj_weight.shape[1]
    w = torch.cat([mha_q.in_proj_weight.detach()[:d], mha_kv.in_proj_weight.detach()[d:]], dim=0)
    b = torch.cat([mha_q.in_proj_bias.detach()[:d], mha_kv.in_proj_bias.detach()[d:]], dim=0)
    return w, b


def prep_mha_self(mha) -> dict:
    return dict(w_qkv=w16(mha.in_proj_weight), b_qkv=v32(mha.in_proj_bias),
                w_o=w16(mha.out_proj.weight), b_o=v32(mha.out_proj.bias))


def prep_cross_pair(mha_q, mha_kv) -> dict:
    """Operands of the fused projection of ONE stream that is the query side of `mha_q`
    and the key/value side of `mha_kv`: rows [Wq(mha_q); Wk(mha_kv); Wv(mha_kv)]."""
    w, b = cross_pair_weights(mha_q, mha_kv)
    return dict(w_qkv=w16(w), b_qkv=v32(b))


def prep_ln(ln) -> Tuple[torch.Tensor, torch.Tensor]:
    return v32(ln.weight), v32(ln.bias)


def prep_linear(lin, k_pad: Optional[int] = None) -> dict:
    return dict(w=w16(lin.weight, k_pad), b=v32(lin.bias) if lin.bias is not None else None)


# --------------------------------------------------------------------------- #
# encoder building blocks
# --------------------------------------------------------------------------- #
# Every sub-layer of the reference is LN(x + f(x)).  With lazy=True the LayerNorm is not run as a
# pass of its own: the out-projection / FFN GEMM writes the pre-LayerNorm sum together with per-row
# statistics, and the result is a Seq carrying a LazyLN that its consumers apply in their epilogues
# (projection GEMMs through folded weights, residual adds, the gate kernels).  lazy=False keeps the
# explicit LayerNorm kernel (legacy block, utterance-level classifier, anything that needs the
# normalised tensor itself).
def materialize(x: Seq, want_f32: bool = False) -> Seq:
    """Apply a pending LayerNorm with the stand-alone kernel."""
    if x.ln is None:
        if want_f32 and x.x32 is None:
            raise L.HriemoError("internal: fp32 copy of a materialised stream was not requested")
        return x
    yb, yf = ops.layernorm(x.x, x.ln.gamma, x.ln.beta, want_bf16=True, want_f32=want_f32)
    return Seq(yb, x.B, x.T, yf)


def residual_ln(x_pre: torch.Tensor, ln, B: int, T: int, want_f32: bool = False) -> Seq:
    yb, yf = ops.layernorm(x_pre, ln[0], ln[1], want_bf16=True, want_f32=want_f32)
    return Seq(yb, B, T, yf)


def project(x: Seq, P: dict, P_folded: Optional[dict], epilogue: int, tag: str) -> torch.Tensor:
    """Linear over the logical value of x: plain operands, or the folded ones when x is lazy."""
    if x.ln is None:
        return ops.gemm(x.x, P["w"], P["b"], epilogue, tag=tag)
    if P_folded is None:
        return project(materialize(x), P, None, epilogue, tag)
    return ops.gemm(x.x, P_folded["w"], P_folded["b"], epilogue, a_ln=(x.ln.stats, P_folded["colsum"]), tag=tag)


def out_residual(o: torch.Tensor, w, b, x: Seq, ln, lazy: bool, tag: str, want_f32: bool = False) -> Seq:
    """LN(x + o W^T + b): the residual is the logical value of x (normalised on the fly if x is lazy)."""
    resid_ln = None if x.ln is None else (x.ln.stats, x.ln.gamma, x.ln.beta)
    if lazy and not want_f32:
        pre, stats = ops.gemm(o, w, b, L.EPI_BIAS_RESID, resid=x.x, resid_ln=resid_ln, want_stats=True, tag=tag)
        return Seq(pre, x.B, x.T, None, LazyLN(ln[0], ln[1], stats))
    pre = ops.gemm(o, w, b, L.EPI_BIAS_RESID, resid=x.x, resid_ln=resid_ln, tag=tag)
    return residual_ln(pre, ln, x.B, x.T, want_f32)


def self_attention_block(x: Seq, P: dict, ln, mask, n_heads: int, want_attn: bool, lazy: bool = False,
                         P_folded: Optional[dict] = None):
    """LN(x + MHA(x, x, x)): models/cross_modal_block_tacfn.py:74-82 / :85-93."""
    d = x.d
    dh = d // n_heads
    if x.ln is not None and P_folded is None:
        x = materialize(x)
    qkv = project(x, dict(w=P["w_qkv"], b=P["b_qkv"]), P_folded, L.EPI_BIAS, "attn_proj")
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    o = ops.attention(q, k, v, mask, x.B, n_heads, x.T, x.T, dh)
    amap = ops.attention_probs(q, k, mask, x.B, n_heads, x.T, x.T, dh) if want_attn else None
    return out_residual(o, P["w_o"], P["b_o"], x, ln, lazy, "attn_proj"), amap


def cross_projection(x: Seq, P: dict, P_folded: Optional[dict] = None):
    """One GEMM producing this stream's cross-attention query, and the key / value it
    offers to the other stream (SURVEY Appendix C "free algebraic fusions")."""
    d = x.d
    qkv = project(x, dict(w=P["w_qkv"], b=P["b_qkv"]), P_folded, L.EPI_BIAS, "attn_proj")
    return qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]


def cross_attention_block(xq: Seq, q, k_other, v_other, T_other: int, mask_other, w_o, b_o, ln,
                          n_heads: int, want_attn: bool, lazy: bool = False):
    """LN(x + MHA(x, other, other)): models/cross_modal_block_tacfn.py:98-105 / :111-118."""
    dh = xq.d // n_heads
    o = ops.attention(q, k_other, v_other, mask_other, xq.B, n_heads, xq.T, T_other, dh)
    amap = ops.attention_probs(q, k_other, mask_other, xq.B, n_heads, xq.T, T_other, dh) if want_attn else None
    return out_residual(o, w_o, b_o, xq, ln, lazy, "attn_proj"), amap


def ffn_block(x: Seq, P1: dict, P2: dict, ln, want_f32: bool = False, lazy: bool = False,
              P1_folded: Optional[dict] = None) -> Seq:
    """LN(x + W2 relu(W1 x + b1) + b2): models/cross_modal_block_tacfn.py:106 / :119."""
    if x.ln is not None and P1_folded is None:
        x = materialize(x)
    h = project(x, P1, P1_folded, L.EPI_BIAS_RELU, "ffn")
    out = out_residual(h, P2["w"], P2["b"], x, ln, lazy, "ffn", want_f32)
    del h
    return out


# --------------------------------------------------------------------------- #
# decoder building blocks (fp32 residual stream + bf16 GEMM operand)
# --------------------------------------------------------------------------- #
def dec_residual_ln(pre32: torch.Tensor, ln) -> Tuple[torch.Tensor, torch.Tensor]:
    return ops.layernorm(pre32, ln[0], ln[1], want_bf16=True, want_f32=True)


def _dec_out_residual(x, w, b, z32, dk):
    """z + dropout(x W^T + b) as fp32 (the decoder's residual stream); dk = (p8, scale, key) | None."""
    if dk is None:
        return ops.gemm(x, w, b, L.EPI_BIAS_RESID_F32, resid=z32)
    return ops.dropout(ops.gemm(x, w, b, L.EPI_BIAS_F32), dk, resid=z32)


def decoder_self_block(zb, z32, P: dict, B: int, Ne: int, n_heads: int, drop=None, site0: int = 0):
    """LN(z + MHA(z, z, z)) over the N_e queries (models/emotion_decoder.py:42-43) -> (zb1, z32_1, (qkv, sa, pre1)).
    drop (training, hriemo.dropout.Drop): site0 + 1 = the attention probabilities, site0 + 2 = dropout1."""
    d = zb.shape[1]
    dh = d // n_heads
    qkv = ops.gemm(zb, P["self"]["w_qkv"], P["self"]["b_qkv"], L.EPI_BIAS)
    sa, _ = ops.small_attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], None, B, n_heads, Ne, Ne, dh,
                                drop=drop.site(site0 + 1) if drop else None)
    pre1 = _dec_out_residual(sa, P["self"]["w_o"], P["self"]["b_o"], z32, drop.site(site0 + 2) if drop else None)
    zb1, z32 = dec_residual_ln(pre1, P["norm1"])
    return zb1, z32, (qkv, sa, pre1)


def decoder_layer(zb, z32, kv_mem, mem_mask, P: dict, B: int, Ne: int, Lm: int, n_heads: int,
                  want_attn: bool, tape: Optional[dict] = None, pre_self=None, drop=None, site0: int = 0):
    """models/emotion_decoder.py:33-64.  kv_mem = [K|V] projection of the memory for this layer.
    tape (training): receives the activations the backward pass needs (hriemo/backward.py).
    pre_self = (zb1, z32_1): the self-attention block's output when it is already known -- the first layer's input is
    the broadcast Parameter (models/emotion_decoder.py:127), so its self-attention block does not depend on the
    utterance and is computed once per weight set (EmotionDecoder._build).
    drop (training, hriemo.dropout.Drop) with the layer's first site id site0: +1 self-attention probabilities, +2 dropout1,
    +3 cross-attention probabilities, +4 dropout2, +5 the FFN's inner dropout, +6 dropout3 (models/emotion_decoder.py:14-29)."""
    d = (zb if pre_self is None else pre_self[0]).shape[1]
    dh = d // n_heads
    zb_in = zb
    if pre_self is None:
        zb1, z32, (qkv, sa, pre1) = decoder_self_block(zb, z32, P, B, Ne, n_heads, drop, site0)
    else:
        zb1, z32 = pre_self
        qkv = sa = pre1 = None
    q = ops.gemm(zb1, P["cross_wq"], P["cross_bq"], L.EPI_BIAS)
    ds = (lambda k: drop.site(site0 + k)) if drop else (lambda k: None)
    ca, probs = ops.small_attention(q, kv_mem[:, :d], kv_mem[:, d:], mem_mask, B, n_heads, Ne, Lm, dh,
                                    want_probs=want_attn, drop=ds(3))
    pre2 = _dec_out_residual(ca, P["cross_wo"], P["cross_bo"], z32, ds(4))
    zb2, z32 = dec_residual_ln(pre2, P["norm2"])
    h = ops.gemm(zb2, P["lin1"]["w"], P["lin1"]["b"], L.EPI_BIAS_RELU)
    if drop:
        h = ops.dropout(h, ds(5))   # the tape keeps the DROPPED hidden: linear2's input, and (h > 0) = kept and active
    pre3 = _dec_out_residual(h, P["lin2"]["w"], P["lin2"]["b"], z32, ds(6))
    zb, z32 = dec_residual_ln(pre3, P["norm3"])
    if tape is not None:
        tape.update(zb_in=zb_in, qkv=qkv, sa=sa, pre1=pre1, zb1=zb1, qc=q, kv_mem=kv_mem, ca=ca, pre2=pre2, zb2=zb2,
                    h=h, pre3=pre3)
    return zb, z32, probs
