"""Backward pass from the loss to the encoder outputs (BASELINE config 5).

What `loss.backward()` does in scripts/fusion/train_fusion_seq_level_decoder.py:332 for everything downstream of
the cross-modal encoder: the loss (:319-327), the emotion decoder (models/emotion_decoder.py:33-64, :117-162) and
the vector beta-gate (models/beta_gate_tacfn.py:68-118).  It is a schedule of C-ABI kernels like engine.py, not an
autograd graph: the training-mode forward keeps the activations the backward needs in plain dicts ("tapes") and
the backward walks the sub-layers in reverse order.

Gradients are returned under the reference's own parameter names (what `named_parameters()` of the reference
model yields), fp32, so that they can be written into the flat gradient arena of train_ops.cu.  Activation
gradients travel as bf16 like the activations; the gate MLP and the head are fp32 in both directions.

The gate's training forward differs from the inference schedule in one respect: LN(a) and LN(t) are written to
HBM as bf16 tensors (the backward reads them twice) instead of being applied on the fly inside the pooling and
blend kernels; streams that arrive with a pending encoder LayerNorm are materialised first.

Dropout (hriemo/dropout.py, csrc/dropout.cuh): every site of the reference -- the sub-layer outputs, the decoder FFN's
inner dropout and the attention probabilities of all six MHAs -- is applied in the training forward when the modules were
built with p > 0 and are in train() mode, with counter-based masks that the backward recomputes from the same stream
keys (a `Drop` object per forward / backward pair; nothing is stored).  With p = 0 (the parity configuration of SURVEY
sec. 8d config 5) the schedules below are exactly the round-1 ones.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import dropout as D
from . import engine as E
from . import lib as L
from . import ops

bf16 = torch.bfloat16
f32 = torch.float32
Grads = Dict[str, torch.Tensor]


def zero_key_bias_gradients(grads: Grads) -> None:
    """The key third of an attention in-projection bias has a gradient that is zero in exact arithmetic: adding a
    constant to every key shifts all scores of a query by the same amount, which softmax ignores.  What arrives there
    is round-off — ~1e-10 in the reference's float32 backward, below AdamW's eps, so the reference barely moves that
    third; ~1e-6 from bf16 activation gradients, above eps, which AdamW would turn into steps of size lr.  The
    round-off is replaced by the exact value."""
    for name, g in grads.items():
        if name.endswith("in_proj_bias"):
            d = g.numel() // 3
            g[d:2 * d].zero_()


def _t(w: torch.Tensor) -> torch.Tensor:
    """[N, K] bf16 weight -> [K, N]: the operand of dX = dY . W for the forward GEMM kernel."""
    return ops.transpose_bf16(w)


# --------------------------------------------------------------------------- #
# beta-gate
# --------------------------------------------------------------------------- #
def gate_forward_train(gate, a: E.Seq, t: E.Seq, mask_a, mask_t) -> Tuple[E.Seq, torch.Tensor, dict]:
    """BetaGate forward (beta_gate_tacfn.py:79-116) keeping what gate_backward needs.  -> (h, beta [B,1], tape)."""
    if a.T < t.T:
        raise RuntimeError(f"BetaGate: audio length {a.T} is shorter than text length {t.T}")
    P = gate._prep.get()
    a, t = E.materialize(a), E.materialize(t)
    na, _ = ops.layernorm(a.x, *P["norm_a"])                                                   # :79
    nt, _ = ops.layernorm(t.x, *P["norm_t"])                                                   # :80
    a_pool = ops.ln_masked_mean(na, None, None, mask_a, a.B, a.T, apply_ln=False)              # :83
    t_pool = ops.ln_masked_mean(nt, None, None, mask_t, t.B, t.T, apply_ln=False)              # :84
    g = ops.gate_input(a_pool, t_pool)                                                         # :87-89
    if "w0_3" in P:   # layer 1 on the split-operand tcgen05 GEMM, as in inference (beta within 1e-7 of the fp32 FMA loop)
        hid = torch.relu(ops.gemm(ops.split3(g), P["w0_3"], P["b0"], L.EPI_BIAS_F32))          # relu keeps NaN
    else:
        hid = ops.sgemm(g, P["w0"], P["b0"], L.ACT_RELU)
    w = ops.sgemm(hid, P["w2"], P["b2"], L.ACT_SIGMOID)                                        # :92
    hb, _, beta = ops.gate_blend(na, a.T, nt, None, None, w, a.B, t.T, apply_ln=False)         # :95-116
    tape = dict(a=a, t=t, na=na, nt=nt, a_pool=a_pool, t_pool=t_pool, g=g, hid=hid, w=w, mask_a=mask_a, mask_t=mask_t)
    return E.Seq(hb, a.B, t.T), beta, tape


def gate_backward(gate, tape: dict, dh: torch.Tensor, dbeta: Optional[torch.Tensor]):
    """dh: bf16 [B*L, d] gradient of the fused sequence, dbeta: fp32 [B, 1] | None.
    -> (d_a bf16 [B*T_a, d], d_t bf16 [B*T_t, d], grads under the names of BetaGate's parameters)."""
    P = gate._prep.get()
    a, t = tape["a"], tape["t"]
    B, T_a, Lf = a.B, a.T, t.T
    G: Grads = {}
    dw = ops.gate_blend_backward_w(dh, tape["na"], T_a, tape["nt"], dbeta, B, Lf)              # :95, :113-116
    d_lin2 = ops.act_backward_f32(dw, tape["w"], L.ACT_SIGMOID)                                # :92
    d_hid, G["mlp.2.weight"], G["mlp.2.bias"] = ops.linear_backward_f32(d_lin2, tape["hid"], P["w2"])
    d_lin0 = ops.act_backward_f32(d_hid, tape["hid"], L.ACT_RELU)
    dg, G["mlp.0.weight"], G["mlp.0.bias"] = ops.linear_backward_f32(d_lin0, tape["g"], P["w0"])
    da_pool, dt_pool = ops.gate_input_backward(dg, tape["a_pool"], tape["t_pool"])             # :87-89
    inv_a = ops.mask_inv_counts(dh, tape["mask_a"], B, T_a)                                    # :20-24
    inv_t = ops.mask_inv_counts(dh, tape["mask_t"], B, Lf)
    d_na = ops.gate_stream_grad(dh, Lf, tape["w"], False, da_pool, tape["mask_a"], inv_a, B, T_a)
    d_nt = ops.gate_stream_grad(dh, Lf, tape["w"], True, dt_pool, tape["mask_t"], inv_t, B, Lf)
    d_a, G["norm_a.weight"], G["norm_a.bias"] = ops.layernorm_backward(a.x, d_na, P["norm_a"][0])   # :79
    d_t, G["norm_t.weight"], G["norm_t.bias"] = ops.layernorm_backward(t.x, d_nt, P["norm_t"][0])   # :80
    return d_a, d_t, G


# --------------------------------------------------------------------------- #
# emotion decoder
# --------------------------------------------------------------------------- #
def decoder_forward_train(dec, mem: E.Seq, mem_mask, drop=None):
    """EmotionDecoder forward (the inference schedule, emotion_decoder.py:117-162) with tapes.
    -> (z [B, N_e, d] fp32, logits [B, N_e] fp32, tape)."""
    if dec.out_proj is None:
        raise L.HriemoError("decoder_forward_train: the training step needs the output layer (use_output_layer=True)")
    tapes: list = []
    z, logits, _ = dec.run(mem, mem_mask, False, tapes=tapes, drop=drop)
    return z, logits, dict(layers=tapes, mem=mem, mem_mask=mem_mask, z=z, drop=drop)


def _decoder_layer_backward(P: dict, tape: dict, dz: torch.Tensor, d_mem: Optional[torch.Tensor], mem: E.Seq, mem_mask,
                            Ne: int, n_heads: int, drop=None, site0: int = 0):
    """Reverse of engine.decoder_layer.  dz: bf16 [B*N_e, d] gradient of the layer's output; d_mem: the gradient of
    the memory accumulated so far (the next layer's share) or None.  -> (dz_in, d_mem, grads)."""
    B, Lm = mem.B, mem.T
    d = dz.shape[1]
    dh = d // n_heads
    dev = dz.device
    G: Grads = {}
    ds = (lambda k: drop.site(site0 + k)) if drop else (lambda k: None)
    dr = (lambda g, k: ops.dropout(g, ds(k))) if drop else (lambda g, k: g)   # the gradient through a dropout site
    # ---- z3 = LN3(z2 + dropout3(W2 dropout(relu(W1 z2 + b1)) + b2))                           :58-59
    d_pre3, G["norm3.weight"], G["norm3.bias"] = ops.layernorm_backward(ops.cast_bf16(tape["pre3"]), dz, P["norm3"][0])
    # tape["h"] is the hidden AFTER the inner dropout: it is linear2's input, and its sign pattern is "active and kept"
    d_hid, G["linear2.weight"], G["linear2.bias"] = ops.linear_backward(dr(d_pre3, 6), tape["h"], _t(P["lin2"]["w"]), relu_input=True)
    d_hid = dr(d_hid, 5)   # the kept elements' 1 / (1 - p) (the mask itself is idempotent)
    G["linear1.weight"], G["linear1.bias"] = ops.linear_wgrad(d_hid, tape["zb2"])
    dz2 = ops.gemm(d_hid, _t(P["lin1"]["w"]), None, L.EPI_BIAS_RESID, resid=d_pre3, tag="dgrad")
    # ---- z2 = LN2(z1 + MHA(z1, mem, mem))                                                     :48-55
    d_pre2, G["norm2.weight"], G["norm2.bias"] = ops.layernorm_backward(ops.cast_bf16(tape["pre2"]), dz2, P["norm2"][0])
    d_ca, G["cross_attn.out_proj.weight"], G["cross_attn.out_proj.bias"] = ops.linear_backward(
        dr(d_pre2, 4), tape["ca"], _t(P["cross_wo"]))
    kv = tape["kv_mem"]
    dq = torch.empty((B * Ne, d), dtype=bf16, device=dev)
    dkv = torch.empty((B * Lm, 2 * d), dtype=bf16, device=dev)
    ops.small_attention_backward(tape["qc"], kv[:, :d], kv[:, d:], d_ca, mem_mask, B, n_heads, Ne, Lm, dh,
                                 out=(dq, dkv[:, :d], dkv[:, d:]), drop=ds(3))
    w_in = torch.empty((3 * d, d), dtype=f32, device=dev)
    b_in = torch.empty((3 * d,), dtype=f32, device=dev)
    ops.linear_wgrad(dq, tape["zb1"], dw=w_in[:d], db=b_in[:d])
    ops.linear_wgrad(dkv, mem.x, dw=w_in[d:], db=b_in[d:])
    G["cross_attn.in_proj_weight"], G["cross_attn.in_proj_bias"] = w_in, b_in
    dz1 = ops.gemm(dq, _t(P["cross_wq"]), None, L.EPI_BIAS_RESID, resid=d_pre2, tag="dgrad")
    if d_mem is None:
        d_mem = ops.gemm(dkv, _t(P["cross_wkv"]), None, L.EPI_BIAS, tag="dgrad")
    else:
        d_mem = ops.gemm(dkv, _t(P["cross_wkv"]), None, L.EPI_BIAS_RESID, resid=d_mem, tag="dgrad")
    # ---- z1 = LN1(z0 + MHA(z0, z0, z0))                                                       :42-43
    d_pre1, G["norm1.weight"], G["norm1.bias"] = ops.layernorm_backward(ops.cast_bf16(tape["pre1"]), dz1, P["norm1"][0])
    d_sa, G["self_attn.out_proj.weight"], G["self_attn.out_proj.bias"] = ops.linear_backward(
        dr(d_pre1, 2), tape["sa"], _t(P["self"]["w_o"]))
    qkv = tape["qkv"]
    dqkv = torch.empty((B * Ne, 3 * d), dtype=bf16, device=dev)
    ops.small_attention_backward(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], d_sa, None, B, n_heads, Ne, Ne, dh,
                                 out=(dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:]), drop=ds(1))
    G["self_attn.in_proj_weight"], G["self_attn.in_proj_bias"] = ops.linear_wgrad(dqkv, tape["zb_in"])
    dz_in = ops.gemm(dqkv, _t(P["self"]["w_qkv"]), None, L.EPI_BIAS_RESID, resid=d_pre1, tag="dgrad")
    return dz_in, d_mem, G


def decoder_backward(dec, tape: dict, d_logits: torch.Tensor, d_z: Optional[torch.Tensor] = None):
    """d_logits: fp32 [B, N_e]; d_z: fp32 [B, N_e, d] | None, the gradient arriving at the decoder's other output z
    (a caller that regularises the emotion embeddings).  -> (d_mem bf16 [B*L, d], grads under the names of
    EmotionDecoder's parameters)."""
    P = dec._prep.get()
    mem, mem_mask = tape["mem"], tape["mem_mask"]
    B, Ne, d = mem.B, dec.num_emotions, dec.d_model
    G: Grads = {}
    # logits = Linear(d, 1)(z).squeeze(-1)                                                      :153-155
    z32 = tape["z"].view(B * Ne, d)
    dz32, G["out_proj.weight"], G["out_proj.bias"] = ops.linear_backward_f32(
        d_logits.contiguous().view(B * Ne, 1), z32, P["w_out"])
    if d_z is not None:
        dz32 = dz32 + d_z.to(f32).reshape(B * Ne, d)
    dz = ops.cast_bf16(dz32)
    d_mem = None
    for i in range(len(dec.layers) - 1, -1, -1):
        layer = dec.layers[i]
        dz, d_mem, g = _decoder_layer_backward(layer._prep.get(), tape["layers"][i], dz, d_mem, mem, mem_mask, Ne,
                                               layer.nhead, tape.get("drop"), 100000 + 100 * i)
        for k, v in g.items():
            G[f"layers.{i}.{k}"] = v
    # the queries are broadcast over the batch (:127): their gradient is the sum over it
    G["emotion_queries"] = ops.sum_rows(dz.view(B, Ne * d)).view(Ne, d)
    if d_mem is None:   # a decoder without layers never looks at the memory
        d_mem = torch.zeros((B * mem.T, d), dtype=bf16, device=dz.device)
    return d_mem, G


# --------------------------------------------------------------------------- #
# cross-modal encoder
# --------------------------------------------------------------------------- #
def _out_residual(o: torch.Tensor, w, b, x: torch.Tensor, dk, tag: str) -> torch.Tensor:
    """x + dropout(o W^T + b): the residual epilogue of the GEMM, or -- with dropout (dk = (p8, scale, key)) -- the plain
    GEMM followed by the dropout + residual pass."""
    if dk is None:
        return ops.gemm(o, w, b, L.EPI_BIAS_RESID, resid=x, tag=tag)
    return ops.dropout(ops.gemm(o, w, b, L.EPI_BIAS, tag=tag), dk, resid=x)


def _ffn_forward(x: torch.Tensor, P1: dict, P2: dict, ln, tape: dict, key: str, dk=None) -> torch.Tensor:
    """LN(x + dropout(W2 relu(W1 x + b1) + b2)) on a materialised bf16 stream (cross_modal_block_tacfn.py:106 / :119)."""
    h = ops.gemm(x, P1["w"], P1["b"], L.EPI_BIAS_RELU, tag="ffn")
    pre = _out_residual(h, P2["w"], P2["b"], x, dk, "ffn")
    y, _ = ops.layernorm(pre, *ln)
    tape[key] = dict(x=x, h=h, pre=pre)
    return y


def _ffn_backward(dy: torch.Tensor, P1: dict, P2: dict, ln, tp: dict, G: Grads, name: str, norm: str, dk=None) -> torch.Tensor:
    d_pre, G[f"{norm}.weight"], G[f"{norm}.bias"] = ops.layernorm_backward(tp["pre"], dy, ln[0])
    d_y = d_pre if dk is None else ops.dropout(d_pre, dk)   # through the sub-layer's dropout; the residual takes d_pre itself
    # d_h = (d_y . W2) * (h > 0): the ReLU mask is the epilogue of the input-gradient GEMM
    d_h, G[f"{name}.2.weight"], G[f"{name}.2.bias"] = ops.linear_backward(d_y, tp["h"], _t(P2["w"]), relu_input=True)
    G[f"{name}.0.weight"], G[f"{name}.0.bias"] = ops.linear_wgrad(d_h, tp["x"])
    return ops.gemm(d_h, _t(P1["w"]), None, L.EPI_BIAS_RESID, resid=d_pre, tag="dgrad")


# dropout sites of encoder layer i: 1000 (i + 1) + kind
_ENC_SITES = dict(self_a_p=1, self_a_o=2, self_t_p=3, self_t_o=4, a2t_p=5, a2t_o=6, ffn_a=7, t2a_p=8, t2a_o=9, ffn_t=10)


def _enc_site(drop, layer: int, kind: str):
    return drop.site(1000 * (layer + 1) + _ENC_SITES[kind]) if drop else None


def encoder_layer_forward_train(block, a: E.Seq, t: E.Seq, mask_a, mask_t, drop=None, layer: int = 0):
    """CrossModalBlock forward (cross_modal_block_tacfn.py:62-125) on materialised streams: every LayerNorm is
    applied by the stand-alone kernel and every sub-layer keeps its input, its pre-LayerNorm sum, the attention
    output with its log-sum-exp and the FFN hidden.  -> (a_out, t_out, tape)."""
    P = block._prep.get()
    H = block.n_heads
    a, t = E.materialize(a), E.materialize(t)
    d = a.d
    dh = d // H
    B, Ta, Tt = a.B, a.T, t.T
    tape: dict = dict(B=B, Ta=Ta, Tt=Tt, mask_a=mask_a, mask_t=mask_t, drop=drop, layer=layer)
    site = lambda kind: _enc_site(drop, layer, kind)

    def self_block(x, Pm, ln, mask, T, key):                                                  # :74-82 / :85-93
        qkv = ops.gemm(x, Pm["w_qkv"], Pm["b_qkv"], L.EPI_BIAS, tag="attn_proj")
        o, lse = ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], mask, B, H, T, T, dh, want_lse=True,
                               drop=site(key + "_p"))
        pre = _out_residual(o, Pm["w_o"], Pm["b_o"], x, site(key + "_o"), "attn_proj")
        y, _ = ops.layernorm(pre, *ln)
        tape[key] = dict(x=x, qkv=qkv, o=o, lse=lse, pre=pre)
        return y

    a_s = self_block(a.x, P["self_a"], P["self_norm_a"], mask_a, Ta, "self_a")
    t_s = self_block(t.x, P["self_t"], P["self_norm_t"], mask_t, Tt, "self_t")
    # one packed projection per stream: its cross-attention query and the key / value it offers the other stream
    qkv_a = ops.gemm(a_s, P["cross_a"]["w_qkv"], P["cross_a"]["b_qkv"], L.EPI_BIAS, tag="attn_proj")
    qkv_t = ops.gemm(t_s, P["cross_t"]["w_qkv"], P["cross_t"]["b_qkv"], L.EPI_BIAS, tag="attn_proj")
    tape.update(a_s=a_s, t_s=t_s, qkv_a=qkv_a, qkv_t=qkv_t)

    def cross_block(x, q, k, v, mask_kv, Tq, Tk, Po, ln, key):                                # :98-105 / :111-118
        o, lse = ops.attention(q, k, v, mask_kv, B, H, Tq, Tk, dh, want_lse=True, drop=site(key + "_p"))
        pre = _out_residual(o, Po["w"], Po["b"], x, site(key + "_o"), "attn_proj")
        y, _ = ops.layernorm(pre, *ln)
        tape[key] = dict(o=o, lse=lse, pre=pre)
        return y

    a1 = cross_block(a_s, qkv_a[:, :d], qkv_t[:, d:2 * d], qkv_t[:, 2 * d:], mask_t, Ta, Tt, P["a2t_o"], P["norm_a1"], "a2t")
    a_o = _ffn_forward(a1, P["ffn_a1"], P["ffn_a2"], P["norm_a2"], tape, "ffn_a", site("ffn_a"))   # :106
    t1 = cross_block(t_s, qkv_t[:, :d], qkv_a[:, d:2 * d], qkv_a[:, 2 * d:], mask_a, Tt, Ta, P["t2a_o"], P["norm_t1"], "t2a")
    t_o = _ffn_forward(t1, P["ffn_t1"], P["ffn_t2"], P["norm_t2"], tape, "ffn_t", site("ffn_t"))   # :119
    return E.Seq(a_o, B, Ta), E.Seq(t_o, B, Tt), tape


def encoder_layer_backward(block, tape: dict, d_a: torch.Tensor, d_t: torch.Tensor, need_dx: bool = True):
    """Reverse of encoder_layer_forward_train.  d_a [B*T_a, d], d_t [B*T_t, d]: bf16 gradients of the layer's outputs.
    -> (d_a_in | None, d_t_in | None, grads under the names of CrossModalBlock's parameters).  need_dx=False skips the
    gradient of the layer's inputs (the first layer reads frozen features)."""
    P = block._prep.get()
    H = block.n_heads
    B, Ta, Tt = tape["B"], tape["Ta"], tape["Tt"]
    mask_a, mask_t = tape["mask_a"], tape["mask_t"]
    d = d_a.shape[1]
    dh = d // H
    dev = d_a.device
    G: Grads = {}
    drop, layer = tape.get("drop"), tape.get("layer", 0)
    site = lambda kind: _enc_site(drop, layer, kind)
    dr = (lambda g, kind: ops.dropout(g, site(kind))) if drop else (lambda g, kind: g)   # the gradient through a dropout site
    qkv_a, qkv_t = tape["qkv_a"], tape["qkv_t"]
    dqkv_a = torch.empty((B * Ta, 3 * d), dtype=bf16, device=dev)
    dqkv_t = torch.empty((B * Tt, 3 * d), dtype=bf16, device=dev)
    # ---- audio: FFN (:106), then a_s + MHA_a2t(a_s, t_s, t_s) (:98-105)
    d_a1 = _ffn_backward(d_a, P["ffn_a1"], P["ffn_a2"], P["norm_a2"], tape["ffn_a"], G, "ffn_a", "norm_a2", site("ffn_a"))
    tp = tape["a2t"]
    d_pre_a1, G["norm_a1.weight"], G["norm_a1.bias"] = ops.layernorm_backward(tp["pre"], d_a1, P["norm_a1"][0])
    d_o, G["attn_a2t.out_proj.weight"], G["attn_a2t.out_proj.bias"] = ops.linear_backward(dr(d_pre_a1, "a2t_o"), tp["o"], _t(P["a2t_o"]["w"]))
    ops.attention_backward(qkv_a[:, :d], qkv_t[:, d:2 * d], qkv_t[:, 2 * d:], tp["o"], d_o, tp["lse"], mask_t, B, H, Ta, Tt, dh,
                           grads=(dqkv_a[:, :d], dqkv_t[:, d:2 * d], dqkv_t[:, 2 * d:]), drop=site("a2t_p"))
    # ---- text: FFN (:119), then t_s + MHA_t2a(t_s, a_s, a_s) (:111-118)
    d_t1 = _ffn_backward(d_t, P["ffn_t1"], P["ffn_t2"], P["norm_t2"], tape["ffn_t"], G, "ffn_t", "norm_t2", site("ffn_t"))
    tp = tape["t2a"]
    d_pre_t1, G["norm_t1.weight"], G["norm_t1.bias"] = ops.layernorm_backward(tp["pre"], d_t1, P["norm_t1"][0])
    d_o, G["attn_t2a.out_proj.weight"], G["attn_t2a.out_proj.bias"] = ops.linear_backward(dr(d_pre_t1, "t2a_o"), tp["o"], _t(P["t2a_o"]["w"]))
    ops.attention_backward(qkv_t[:, :d], qkv_a[:, d:2 * d], qkv_a[:, 2 * d:], tp["o"], d_o, tp["lse"], mask_a, B, H, Tt, Ta, dh,
                           grads=(dqkv_t[:, :d], dqkv_a[:, d:2 * d], dqkv_a[:, 2 * d:]), drop=site("t2a_p"))
    # ---- the packed projections: rows [Wq(a2t); Wk(t2a); Wv(t2a)] read a_s, rows [Wq(t2a); Wk(a2t); Wv(a2t)] read t_s
    for n in ("attn_a2t", "attn_t2a"):
        G[f"{n}.in_proj_weight"] = torch.empty((3 * d, d), dtype=f32, device=dev)
        G[f"{n}.in_proj_bias"] = torch.empty((3 * d,), dtype=f32, device=dev)
    ops.linear_wgrad(dqkv_a[:, :d], tape["a_s"], dw=G["attn_a2t.in_proj_weight"][:d], db=G["attn_a2t.in_proj_bias"][:d])
    ops.linear_wgrad(dqkv_a[:, d:], tape["a_s"], dw=G["attn_t2a.in_proj_weight"][d:], db=G["attn_t2a.in_proj_bias"][d:])
    ops.linear_wgrad(dqkv_t[:, :d], tape["t_s"], dw=G["attn_t2a.in_proj_weight"][:d], db=G["attn_t2a.in_proj_bias"][:d])
    ops.linear_wgrad(dqkv_t[:, d:], tape["t_s"], dw=G["attn_a2t.in_proj_weight"][d:], db=G["attn_a2t.in_proj_bias"][d:])
    d_a_s = ops.gemm(dqkv_a, _t(P["cross_a"]["w_qkv"]), None, L.EPI_BIAS_RESID, resid=d_pre_a1, tag="dgrad")
    d_t_s = ops.gemm(dqkv_t, _t(P["cross_t"]["w_qkv"]), None, L.EPI_BIAS_RESID, resid=d_pre_t1, tag="dgrad")

    def self_backward(dy, Pm, ln, mask, T, tp, name, norm, buf, kind):                         # :74-82 / :85-93
        d_pre, G[f"{norm}.weight"], G[f"{norm}.bias"] = ops.layernorm_backward(tp["pre"], dy, ln[0])
        d_o, G[f"{name}.out_proj.weight"], G[f"{name}.out_proj.bias"] = ops.linear_backward(dr(d_pre, kind + "_o"), tp["o"], _t(Pm["w_o"]))
        qkv = tp["qkv"]
        ops.attention_backward(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], tp["o"], d_o, tp["lse"], mask, B, H, T, T, dh,
                               grads=(buf[:, :d], buf[:, d:2 * d], buf[:, 2 * d:]), drop=site(kind + "_p"))
        G[f"{name}.in_proj_weight"], G[f"{name}.in_proj_bias"] = ops.linear_wgrad(buf, tp["x"])
        if not need_dx:
            return None
        return ops.gemm(buf, _t(Pm["w_qkv"]), None, L.EPI_BIAS_RESID, resid=d_pre, tag="dgrad")

    # the packed cross-projection gradients are consumed: their buffers are reused for the self-attention ones
    d_a_in = self_backward(d_a_s, P["self_a"], P["self_norm_a"], mask_a, Ta, tape["self_a"], "self_attn_a", "self_norm_a", dqkv_a, "self_a")
    d_t_in = self_backward(d_t_s, P["self_t"], P["self_norm_t"], mask_t, Tt, tape["self_t"], "self_attn_t", "self_norm_t", dqkv_t, "self_t")
    return d_a_in, d_t_in, G


def encoder_forward_train(enc, a: E.Seq, t: E.Seq, mask_a, mask_t, drop=None):
    """CrossModalTransformer forward (cross_modal_block_tacfn.py:145-166) with one tape per layer."""
    tapes = []
    for i, block in enumerate(enc.layers):
        a, t, tape = encoder_layer_forward_train(block, a, t, mask_a, mask_t, drop, i)
        tapes.append(tape)
    return a, t, tapes


def encoder_backward(enc, tapes: list, d_a: torch.Tensor, d_t: torch.Tensor, need_dx: bool = False, before_first_layer=None):
    """-> (d_a_in | None, d_t_in | None, grads under the names of CrossModalTransformer's parameters).
    before_first_layer(G): called once, just before the backward of layer 0 (the last one to run) starts, with the
    gradients of layers >= 1 -- the point from which a data-parallel trainer can exchange everything but layer 0."""
    G: Grads = {}
    for i in range(len(enc.layers) - 1, -1, -1):
        if i == 0 and before_first_layer is not None:
            before_first_layer(G)
        d_a, d_t, g = encoder_layer_backward(enc.layers[i], tapes[i], d_a, d_t, need_dx=need_dx or i > 0)
        for k, v in g.items():
            G[f"layers.{i}.{k}"] = v
    return d_a, d_t, G


# --------------------------------------------------------------------------- #
# loss -> encoder outputs
# --------------------------------------------------------------------------- #
def drop_for(model) -> Optional[D.Drop]:
    """The dropout of one training pass: the model's own rate while it is in train() mode (a fresh seed per call from
    torch's CPU generator), None in eval() mode or for p = 0."""
    return D.make(float(getattr(model, "p_drop", 0.0))) if model.training else None


def decode_loss_and_backward(model, a: E.Seq, t: E.Seq, mask_a, mask_t, labels: torch.Tensor,
                             beta_weight: float = 0.01, drop=None) -> dict:
    """Gate -> decoder -> loss and back, for a FusionWithEmotionDecoder-like `model` (attributes beta_gate and
    emotion_decoder) given the encoder outputs a / t.  Returns a dict:
      loss [1], logits [B, N_e], beta [B, 1], z [B, N_e, d]  (fp32);
      grads: {"beta_gate.*", "emotion_decoder.*": fp32 tensors shaped like the parameters};
      d_a [B*T_a, d], d_t [B*T_t, d]: bf16 gradients of the encoder outputs (input of the encoder's backward)."""
    h, beta, gate_tape = gate_forward_train(model.beta_gate, a, t, mask_a, mask_t)
    fused_mask = model._build_fused_mask(mask_a, mask_t, h.T)
    z, logits, dec_tape = decoder_forward_train(model.emotion_decoder, h, fused_mask, drop)
    loss, d_logits, d_beta = ops.bce_beta_loss(logits, labels, beta, beta_weight)
    d_h, g_dec = decoder_backward(model.emotion_decoder, dec_tape, d_logits)
    d_a, d_t, g_gate = gate_backward(model.beta_gate, gate_tape, d_h, d_beta)
    grads: Grads = {f"beta_gate.{k}": v for k, v in g_gate.items()}
    grads.update({f"emotion_decoder.{k}": v for k, v in g_dec.items()})
    zero_key_bias_gradients(grads)
    return dict(loss=loss, logits=logits, beta=beta, z=z, grads=grads, d_a=d_a, d_t=d_t)


def loss_and_gradients(model, h_a: torch.Tensor, h_t: torch.Tensor, mask_a, mask_t, labels: torch.Tensor,
                       beta_weight: float = 0.01, early_hook=None) -> dict:
    """Forward and backward of one training iteration of FusionWithEmotionDecoder
    (scripts/fusion/train_fusion_seq_level_decoder.py:311-332: model(...), BCE + beta regulariser, loss.backward()).
    h_a [B, T_a, d], h_t [B, T_t, d] fp32 / bf16 CUDA features, masks bool True = PAD, labels [B, N_e].
    -> dict(loss, logits, beta, z, grads): grads holds one fp32 tensor per parameter of the model, under the
    reference's parameter names.
    early_hook(grads_so_far): called once, before the backward of encoder layer 0 starts, with the final gradients of
    every other parameter (decoder, gate, encoder layers >= 1) under their full names: Trainer starts their
    all-reduce there, so that it runs under the last layer's backward."""
    a, t = E.to_seq(h_a, "h_a"), E.to_seq(h_t, "h_t")
    mask_a = E.check_mask(mask_a, a.B, a.T, "mask_a")
    mask_t = E.check_mask(mask_t, t.B, t.T, "mask_t")
    drop = drop_for(model)
    a_enc, t_enc, enc_tapes = encoder_forward_train(model.cross_modal, a, t, mask_a, mask_t, drop)
    out = decode_loss_and_backward(model, a_enc, t_enc, mask_a, mask_t, labels, beta_weight, drop)
    hook = None
    if early_hook is not None:
        def hook(g_layers):
            early = dict(out["grads"])
            early.update({f"cross_modal.{k}": v for k, v in g_layers.items()})
            zero_key_bias_gradients(early)
            early_hook(early)
    _, _, g_enc = encoder_backward(model.cross_modal, enc_tapes, out.pop("d_a"), out.pop("d_t"), before_first_layer=hook)
    out["grads"].update({f"cross_modal.{k}": v for k, v in g_enc.items()})
    zero_key_bias_gradients(out["grads"])
    return out


# --------------------------------------------------------------------------- #
# forward with tapes / backward from output gradients: the autograd boundary (hriemo/autograd.py)
# --------------------------------------------------------------------------- #
def forward_train(model, h_a: torch.Tensor, h_t: torch.Tensor, mask_a, mask_t):
    """The training-mode forward of FusionWithEmotionDecoder (models/fusion_with_emotion_decoder.py:120-197) keeping the
    tapes of every sub-layer.  -> (logits [B, N_e], beta [B, 1], z [B, N_e, d], ctx) with ctx what backward_from needs."""
    a, t = E.to_seq(h_a, "h_a"), E.to_seq(h_t, "h_t")
    mask_a = E.check_mask(mask_a, a.B, a.T, "mask_a")
    mask_t = E.check_mask(mask_t, t.B, t.T, "mask_t")
    drop = drop_for(model)
    a_enc, t_enc, enc_tapes = encoder_forward_train(model.cross_modal, a, t, mask_a, mask_t, drop)
    h, beta, gate_tape = gate_forward_train(model.beta_gate, a_enc, t_enc, mask_a, mask_t)
    fused_mask = model._build_fused_mask(mask_a, mask_t, h.T)
    z, logits, dec_tape = decoder_forward_train(model.emotion_decoder, h, fused_mask, drop)
    return logits, beta, z, dict(enc=enc_tapes, gate=gate_tape, dec=dec_tape)


def backward_from(model, ctx: dict, d_logits: Optional[torch.Tensor], d_beta: Optional[torch.Tensor],
                  d_z: Optional[torch.Tensor] = None, need_dx: bool = False):
    """What loss.backward() does from the gradients of the model's three outputs (any of them may be None): decoder,
    gate and encoder backward schedules.  -> (grads under the reference's parameter names, d_a_in, d_t_in) -- the
    last two (bf16 [B*T, d] gradients of the model's inputs) only with need_dx (the MOSEI wrapper's projections)."""
    dec = model.emotion_decoder
    B, Ne = ctx["dec"]["mem"].B, dec.num_emotions
    dev = ctx["dec"]["z"].device
    if d_logits is None:
        d_logits = torch.zeros((B, Ne), dtype=f32, device=dev)
    d_h, g_dec = decoder_backward(dec, ctx["dec"], d_logits.to(f32).contiguous(), d_z)
    d_a, d_t, g_gate = gate_backward(model.beta_gate, ctx["gate"], d_h, None if d_beta is None else d_beta.to(f32).contiguous())
    d_a_in, d_t_in, g_enc = encoder_backward(model.cross_modal, ctx["enc"], d_a, d_t, need_dx=need_dx)
    grads: Grads = {f"beta_gate.{k}": v for k, v in g_gate.items()}
    grads.update({f"emotion_decoder.{k}": v for k, v in g_dec.items()})
    grads.update({f"cross_modal.{k}": v for k, v in g_enc.items()})
    zero_key_bias_gradients(grads)
    return grads, d_a_in, d_t_in
