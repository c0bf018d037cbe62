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
class Seq:
    """A batch of sequences as one row-major matrix: x[b*T + t, :]."""
    x: torch.Tensor                      # bf16 [B*T, d]
    B: int
    T: int
    x32: Optional[torch.Tensor] = None   # fp32 copy when the caller asked for one

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
    global _warned_train
    if module.training and p_drop > 0 and not _warned_train:
        _warned_train = True
        warnings.warn(
            "hri-emo_b200: forward-only build — dropout is not applied and no autograd graph is "
            "recorded even though the module is in training mode; call .eval() for inference.",
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


def prep_mha_self(mha) -> dict:
    return dict(w_qkv=w16(mha.in_proj_weight), b_qkv=v32(mha.in_proj_bias),
                w_o=w16(mha.out_proj.weight), b_o=v32(mha.out_proj.bias))


def prep_cross_pair(mha_q, mha_kv) -> dict:
    """Operands of the fused projection of ONE stream that is the query side of `mha_q`
    and the key/value side of `mha_kv`: rows [Wq(mha_q); Wk(mha_kv); Wv(mha_kv)]."""
    d = mha_q.in_proj_weight.shape[1]
    w = torch.cat([mha_q.in_proj_weight.detach()[:d], mha_kv.in_proj_weight.detach()[d:]], dim=0)
    b = torch.cat([mha_q.in_proj_bias.detach()[:d], mha_kv.in_proj_bias.detach()[d:]], dim=0)
    return dict(w_qkv=w16(w), b_qkv=v32(b))


def prep_ln(ln) -> Tuple[torch.Tensor, torch.Tensor]:
    return v32(ln.weight), v32(ln.bias)


def prep_linear(lin, k_pad: Optional[int] = None) -> dict:
    return dict(w=w16(lin.weight, k_pad), b=v32(lin.bias) if lin.bias is not None else None)


# --------------------------------------------------------------------------- #
# encoder building blocks
# --------------------------------------------------------------------------- #
def residual_ln(x_pre: torch.Tensor, ln, B: int, T: int, want_f32: bool = False) -> Seq:
    yb, yf = ops.layernorm(x_pre, ln[0], ln[1], want_bf16=True, want_f32=want_f32)
    return Seq(yb, B, T, yf)


def self_attention_block(x: Seq, P: dict, ln, mask, n_heads: int, want_attn: bool):
    """LN(x + MHA(x, x, x)): models/cross_modal_block_tacfn.py:74-82 / :85-93."""
    d = x.d
    dh = d // n_heads
    qkv = ops.gemm(x.x, P["w_qkv"], P["b_qkv"], L.EPI_BIAS)
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    o = ops.attention(q, k, v, mask, x.B, n_heads, x.T, x.T, dh)
    pre = ops.gemm(o, P["w_o"], P["b_o"], L.EPI_BIAS_RESID, resid=x.x)
    amap = ops.attention_probs(q, k, mask, x.B, n_heads, x.T, x.T, dh) if want_attn else None
    return residual_ln(pre, ln, x.B, x.T), amap


def cross_projection(x: Seq, P: dict):
    """One GEMM producing this stream's cross-attention query, and the key / value it
    offers to the other stream (SURVEY Appendix C "free algebraic fusions")."""
    d = x.d
    qkv = ops.gemm(x.x, P["w_qkv"], P["b_qkv"], L.EPI_BIAS)
    return qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]


def cross_attention_block(xq: Seq, q, k_other, v_other, T_other: int, mask_other, w_o, b_o, ln,
                          n_heads: int, want_attn: bool):
    """LN(x + MHA(x, other, other)): models/cross_modal_block_tacfn.py:98-105 / :111-118."""
    dh = xq.d // n_heads
    o = ops.attention(q, k_other, v_other, mask_other, xq.B, n_heads, xq.T, T_other, dh)
    pre = ops.gemm(o, w_o, b_o, L.EPI_BIAS_RESID, resid=xq.x)
    amap = ops.attention_probs(q, k_other, mask_other, xq.B, n_heads, xq.T, T_other, dh) if want_attn else None
    return residual_ln(pre, ln, xq.B, xq.T), amap


def ffn_block(x: Seq, P1: dict, P2: dict, ln, want_f32: bool = False) -> Seq:
    """LN(x + W2 relu(W1 x + b1) + b2): models/cross_modal_block_tacfn.py:106 / :119."""
    h = ops.gemm(x.x, P1["w"], P1["b"], L.EPI_BIAS_RELU)
    pre = ops.gemm(h, P2["w"], P2["b"], L.EPI_BIAS_RESID, resid=x.x)
    del h
    return residual_ln(pre, ln, x.B, x.T, want_f32)


# --------------------------------------------------------------------------- #
# decoder building blocks (fp32 residual stream + bf16 GEMM operand)
# --------------------------------------------------------------------------- #
def dec_residual_ln(pre32: torch.Tensor, ln) -> Tuple[torch.Tensor, torch.Tensor]:
    return ops.layernorm(pre32, ln[0], ln[1], want_bf16=True, want_f32=True)


def decoder_layer(zb, z32, kv_mem, mem_mask, P: dict, B: int, Ne: int, Lm: int, n_heads: int,
                  want_attn: bool):
    """models/emotion_decoder.py:33-64.  kv_mem = [K|V] projection of the memory for this layer."""
    d = zb.shape[1]
    dh = d // n_heads
    qkv = ops.gemm(zb, P["self"]["w_qkv"], P["self"]["b_qkv"], L.EPI_BIAS)
    sa, _ = ops.small_attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], None, B, n_heads, Ne, Ne, dh)
    pre = ops.gemm(sa, P["self"]["w_o"], P["self"]["b_o"], L.EPI_BIAS_RESID_F32, resid=z32)
    zb, z32 = dec_residual_ln(pre, P["norm1"])
    q = ops.gemm(zb, P["cross_wq"], P["cross_bq"], L.EPI_BIAS)
    ca, probs = ops.small_attention(q, kv_mem[:, :d], kv_mem[:, d:], mem_mask, B, n_heads, Ne, Lm, dh,
                                    want_probs=want_attn)
    pre = ops.gemm(ca, P["cross_wo"], P["cross_bo"], L.EPI_BIAS_RESID_F32, resid=z32)
    zb, z32 = dec_residual_ln(pre, P["norm2"])
    h = ops.gemm(zb, P["lin1"]["w"], P["lin1"]["b"], L.EPI_BIAS_RELU)
    pre = ops.gemm(h, P["lin2"]["w"], P["lin2"]["b"], L.EPI_BIAS_RESID_F32, resid=z32)
    zb, z32 = dec_residual_ln(pre, P["norm3"])
    return zb, z32, probs
