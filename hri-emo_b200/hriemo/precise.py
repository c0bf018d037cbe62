""""tf32-class" precision mode of the fusion-and-decode forward (north_star: logits within 1e-4 of the reference's
fp32 forward; the bf16 path's bar is 1e-2).

Every activation stays fp32 in HBM.  A Linear runs on the tcgen05 GEMM of the bf16 path with both operands split into
two bf16 numbers (hi + lo, 16 mantissa bits -- TF32 keeps 10) and three of the four partial products laid out along K
(ops.split3; include/hriemo.h), accumulated in fp32 with an fp32 epilogue.  Softmax attention, LayerNorm, the gate and
the emotion head are fp32 kernels.  Measured against the float64 oracle: logits 5e-6 (CPU emulation of the split) -- the
mode trades throughput for digits and is selected explicitly:

    from hriemo import precise
    with precise.mode("tf32x3"):
        logits, beta, z = model(h_a, h_t, mask_a, mask_t)

or process-wide with HRIEMO_PRECISION=tf32x3 / precise.set_mode("tf32x3").  Inference only (eval / no_grad).

Follows models/cross_modal_block_tacfn.py:62-125, models/beta_gate_tacfn.py:68-118,
models/fusion_with_emotion_decoder.py:120-197, models/emotion_decoder.py:33-64, 117-162 and
models/mosei_fusion_with_emotion_decoder.py:64-66 of the reference, sub-layer by sub-layer.
"""
from __future__ import annotations

import contextlib
import os
from typing import Optional

import torch

from . import engine as E
from . import lib as L
from . import ops

MODES = ("bf16", "tf32x3")
_mode = os.environ.get("HRIEMO_PRECISION", "bf16")
if _mode not in MODES:
    raise L.HriemoError(f"HRIEMO_PRECISION={_mode!r}: expected one of {MODES}")

# utterances are processed in slabs of at most this many audio rows: the widest fp32 intermediate (the 4d FFN hidden
# and its three-way split) stays below ~4 GB
MAX_ROWS_PER_SLAB = 1 << 17


def get_mode() -> str:
    return _mode


def set_mode(mode: str) -> None:
    global _mode
    if mode not in MODES:
        raise L.HriemoError(f"precision mode {mode!r}: expected one of {MODES}")
    _mode = mode


@contextlib.contextmanager
def mode(m: str):
    prev = get_mode()
    set_mode(m)
    try:
        yield
    finally:
        set_mode(prev)


# --------------------------------------------------------------------------- #
# prepared operands: weights split once per parameter version
# --------------------------------------------------------------------------- #
def _w3(w: torch.Tensor) -> torch.Tensor:
    return ops.split3(E.v32(w), weight=True)


def _mha(m) -> dict:
    return dict(w=_w3(m.in_proj_weight), b=E.v32(m.in_proj_bias), wo=_w3(m.out_proj.weight), bo=E.v32(m.out_proj.bias))


def _lin(l) -> dict:
    return dict(w=_w3(l.weight), b=E.v32(l.bias) if l.bias is not None else None)


def _build(model) -> dict:
    enc = []
    for blk in model.cross_modal.layers:
        enc.append(dict(
            self_a=_mha(blk.self_attn_a), self_t=_mha(blk.self_attn_t), a2t=_mha(blk.attn_a2t), t2a=_mha(blk.attn_t2a),
            ffn_a1=_lin(blk.ffn_a[0]), ffn_a2=_lin(blk.ffn_a[2]), ffn_t1=_lin(blk.ffn_t[0]), ffn_t2=_lin(blk.ffn_t[2]),
            **{n: E.prep_ln(getattr(blk, n)) for n in ("self_norm_a", "self_norm_t", "norm_a1", "norm_a2", "norm_t1", "norm_t2")}))
    g = model.beta_gate
    gate = dict(norm_a=E.prep_ln(g.norm_a), norm_t=E.prep_ln(g.norm_t), w0=E.v32(g.mlp[0].weight), b0=E.v32(g.mlp[0].bias),
                w2=E.v32(g.mlp[2].weight), b2=E.v32(g.mlp[2].bias))
    dec = []
    for lay in model.emotion_decoder.layers:
        dec.append(dict(self=_mha(lay.self_attn), cross=_mha(lay.cross_attn), lin1=_lin(lay.linear1), lin2=_lin(lay.linear2),
                        norm1=E.prep_ln(lay.norm1), norm2=E.prep_ln(lay.norm2), norm3=E.prep_ln(lay.norm3)))
    d = model.emotion_decoder
    head = None if d.out_proj is None else (E.v32(d.out_proj.weight), E.v32(d.out_proj.bias))
    return dict(enc=enc, gate=gate, dec=dec, q32=E.v32(d.emotion_queries), head=head)


def _prepared(model) -> dict:
    prep = model.__dict__.get("_precise_prep")
    if prep is None:
        prep = E.Prepared(model, lambda: _build(model))
        model.__dict__["_precise_prep"] = prep
    return prep.get()


# --------------------------------------------------------------------------- #
# sub-layers (fp32 [B*T, d] streams)
# --------------------------------------------------------------------------- #
def _linear(x3: torch.Tensor, w3: torch.Tensor, b, resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    if resid is not None:
        return ops.gemm(x3, w3, b, L.EPI_BIAS_RESID_F32, resid=resid, tag="precise")
    return ops.gemm(x3, w3, b, L.EPI_BIAS_F32, tag="precise")


def _ln(x: torch.Tensor, ln) -> torch.Tensor:
    return ops.layernorm(x, ln[0], ln[1], want_bf16=False, want_f32=True)[1]


def _mha_block(x, x3, kv3, P, ln, mask_kv, B, Tq, Tk, H, want_attn):
    """LN(x + MHA(x, kv, kv)): x3 / kv3 are the split query-side / key-value-side inputs (the same tensor for
    self-attention, where one GEMM forms the packed [Q|K|V])."""
    d = x.shape[1]
    dh = d // H
    if kv3 is x3:
        qkv = _linear(x3, P["w"], P["b"])
        q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    else:
        q = _linear(x3, P["w"][:d], P["b"][:d])
        kv = _linear(kv3, P["w"][d:], P["b"][d:])
        k, v = kv[:, :d], kv[:, d:]
    o, probs = ops.attention_f32(q, k, v, mask_kv, B, H, Tq, Tk, dh, want_probs=want_attn)
    pre = _linear(ops.split3(o), P["wo"], P["bo"], resid=x)
    return _ln(pre, ln), probs


def _ffn_block(x, P1, P2, ln):
    h = _linear(ops.split3(x), P1["w"], P1["b"])
    pre = _linear(ops.split3(h, relu=True), P2["w"], P2["b"], resid=x)
    del h
    return _ln(pre, ln)


def _encoder_layer(P, a, t, mask_a, mask_t, B, T_a, T_t, H, want_attn):
    maps = {}
    a3, t3 = ops.split3(a), ops.split3(t)
    a_s, maps["audio_self"] = _mha_block(a, a3, a3, P["self_a"], P["self_norm_a"], mask_a, B, T_a, T_a, H, want_attn)
    t_s, maps["text_self"] = _mha_block(t, t3, t3, P["self_t"], P["self_norm_t"], mask_t, B, T_t, T_t, H, want_attn)
    del a3, t3
    as3, ts3 = ops.split3(a_s), ops.split3(t_s)
    a1, maps["audio_queries_text"] = _mha_block(a_s, as3, ts3, P["a2t"], P["norm_a1"], mask_t, B, T_a, T_t, H, want_attn)
    t1, maps["text_queries_audio"] = _mha_block(t_s, ts3, as3, P["t2a"], P["norm_t1"], mask_a, B, T_t, T_a, H, want_attn)
    del as3, ts3
    a_o = _ffn_block(a1, P["ffn_a1"], P["ffn_a2"], P["norm_a2"])
    t_o = _ffn_block(t1, P["ffn_t1"], P["ffn_t2"], P["norm_t2"])
    return a_o, t_o, (maps if want_attn else None)


def _gate(P, a, t, mask_a, mask_t, B, T_a, T_t):
    if T_a < T_t:
        raise RuntimeError(f"BetaGate: audio length {T_a} is shorter than text length {T_t}")
    na, nt = _ln(a, P["norm_a"]), _ln(t, P["norm_t"])
    g = ops.gate_input(ops.masked_mean_f32(na, mask_a, B, T_a), ops.masked_mean_f32(nt, mask_t, B, T_t))
    w = ops.sgemm(ops.sgemm(g, P["w0"], P["b0"], L.ACT_RELU), P["w2"], P["b2"], L.ACT_SIGMOID)
    return ops.gate_blend_f32(na, T_a, nt, w, B, T_t)


def _decoder(Pd, q32, head, mem, mem_mask, B, Lm, H, want_attn):
    Ne, d = q32.shape
    z = q32.unsqueeze(0).expand(B, Ne, d).contiguous().view(B * Ne, d)
    mem3 = ops.split3(mem)
    attn = []
    for P in Pd:
        z3 = ops.split3(z)
        z, _ = _mha_block(z, z3, z3, P["self"], P["norm1"], None, B, Ne, Ne, H, False)
        z, probs = _mha_block(z, ops.split3(z), mem3, P["cross"], P["norm2"], mem_mask, B, Ne, Lm, H, want_attn)
        z = _ffn_block(z, P["lin1"], P["lin2"], P["norm3"])
        if want_attn:
            attn.append(probs)
    logits = None if head is None else ops.sgemm(z, head[0], head[1], L.ACT_NONE).view(B, Ne)
    return z.view(B, Ne, d), logits, (attn if want_attn else None)


def _run_slab(model, P, a, t, mask_a, mask_t, B, T_a, T_t, want_attn):
    H = model.cross_modal.layers[0].n_heads if len(model.cross_modal.layers) else model.emotion_decoder.layers[0].nhead
    enc_maps = []
    for Pl in P["enc"]:
        a, t, maps = _encoder_layer(Pl, a, t, mask_a, mask_t, B, T_a, T_t, H, want_attn)
        if want_attn:
            enc_maps.append(maps)
    h, beta = _gate(P["gate"], a, t, mask_a, mask_t, B, T_a, T_t)
    fused_mask = model._build_fused_mask(mask_a, mask_t, T_t)
    Hd = model.emotion_decoder.layers[0].nhead if len(model.emotion_decoder.layers) else H
    z, logits, dec_maps = _decoder(P["dec"], P["q32"], P["head"], h, fused_mask, B, T_t, Hd, want_attn)
    return logits, beta, z, ({"encoder": enc_maps, "decoder": dec_maps} if want_attn else None)


@torch.no_grad()
def fusion_forward(model, h_a, h_t, mask_a, mask_t, want_attn: bool = False, pre=None):
    """FusionWithEmotionDecoder.forward in the tf32x3 mode.  h_a / h_t: [B, T, d] CUDA tensors (already 3-D, masks
    already checked).  pre (MOSEI wrapper): maps the fp32 [rows, d_in] slab inputs to the backbone's streams."""
    P = _prepared(model)
    B, T_a, T_t = h_a.shape[0], h_a.shape[1], h_t.shape[1]
    per = max(1, MAX_ROWS_PER_SLAB // max(T_a, 1))
    outs = []
    for s in range(0, B, per):
        e = min(B, s + per)
        a = h_a[s:e].float().contiguous().view((e - s) * T_a, -1)
        t = h_t[s:e].float().contiguous().view((e - s) * T_t, -1)
        if pre is not None:
            a, t = pre(a, t)
        ma = None if mask_a is None else mask_a[s:e].contiguous()
        mt = None if mask_t is None else mask_t[s:e].contiguous()
        outs.append(_run_slab(model, P, a, t, ma, mt, e - s, T_a, T_t, want_attn))
    if len(outs) == 1:
        return outs[0]
    logits = None if outs[0][0] is None else torch.cat([o[0] for o in outs], dim=0)
    beta = torch.cat([o[1] for o in outs], dim=0)
    z = torch.cat([o[2] for o in outs], dim=0)
    pack = None
    if want_attn:
        enc = [{k: torch.cat([o[3]["encoder"][i][k] for o in outs], dim=0) for k in outs[0][3]["encoder"][i]}
               for i in range(len(outs[0][3]["encoder"]))]
        dec = [torch.cat([o[3]["decoder"][i] for o in outs], dim=0) for i in range(len(outs[0][3]["decoder"]))]
        pack = {"encoder": enc, "decoder": dec}
    return logits, beta, z, pack


@torch.no_grad()
def mosei_forward(wrapper, h_a, h_t, mask_a, mask_t, want_attn: bool = False):
    """MoseiFusionWithEmotionDecoder.forward in the tf32x3 mode (audio_proj / text_proj through the split GEMM)."""
    prep = wrapper.__dict__.get("_precise_prep")
    if prep is None:
        prep = E.Prepared(wrapper, lambda: dict(a=_lin(wrapper.audio_proj), t=_lin(wrapper.text_proj)))
        wrapper.__dict__["_precise_prep"] = prep
    Pw = prep.get()

    def project(a, t):
        return (_linear(ops.split3(a), Pw["a"]["w"], Pw["a"]["b"]), _linear(ops.split3(t), Pw["t"]["w"], Pw["t"]["b"]))

    return fusion_forward(wrapper.backbone, h_a, h_t, mask_a, mask_t, want_attn, pre=project)
