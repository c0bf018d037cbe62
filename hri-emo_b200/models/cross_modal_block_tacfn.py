"""TACFN-style cross-modal encoder — drop-in for the reference's
models/cross_modal_block_tacfn.py (CrossModalBlock :6-127, CrossModalTransformer :130-166)."""
from __future__ import annotations

import torch
import torch.nn as nn

from hriemo import engine as E

from ._containers import MHAParams, ffn


class CrossModalBlock(nn.Module):
    """Intra-modal self-attention, bidirectional audio<->text cross-attention, FFN, six
    post-LayerNorms (reference :62-125).  Parameters keep the reference names."""

    def __init__(self, d_model=768, n_heads=8, dropout=0.1):
        super().__init__()
        self.d_model = d_model
        self.n_heads = n_heads
        self.p_drop = dropout
        self.self_attn_a = MHAParams(d_model, n_heads, dropout)
        self.self_attn_t = MHAParams(d_model, n_heads, dropout)
        self.self_norm_a = nn.LayerNorm(d_model)
        self.self_norm_t = nn.LayerNorm(d_model)
        self.attn_a2t = MHAParams(d_model, n_heads, dropout)
        self.attn_t2a = MHAParams(d_model, n_heads, dropout)
        self.ffn_a = ffn(d_model, 4 * d_model)
        self.ffn_t = ffn(d_model, 4 * d_model)
        self.norm_a1 = nn.LayerNorm(d_model)
        self.norm_a2 = nn.LayerNorm(d_model)
        self.norm_t1 = nn.LayerNorm(d_model)
        self.norm_t2 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self._prep = E.Prepared(self, self._build)

    def _build(self) -> dict:
        return dict(
            self_a=E.prep_mha_self(self.self_attn_a), self_t=E.prep_mha_self(self.self_attn_t),
            # audio stream: query of a2t, key/value of t2a.  text stream: the mirror.
            cross_a=E.prep_cross_pair(self.attn_a2t, self.attn_t2a),
            cross_t=E.prep_cross_pair(self.attn_t2a, self.attn_a2t),
            a2t_o=E.prep_linear(self.attn_a2t.out_proj), t2a_o=E.prep_linear(self.attn_t2a.out_proj),
            ffn_a1=E.prep_linear(self.ffn_a[0]), ffn_a2=E.prep_linear(self.ffn_a[2]),
            ffn_t1=E.prep_linear(self.ffn_t[0]), ffn_t2=E.prep_linear(self.ffn_t[2]),
            self_norm_a=E.prep_ln(self.self_norm_a), self_norm_t=E.prep_ln(self.self_norm_t),
            norm_a1=E.prep_ln(self.norm_a1), norm_a2=E.prep_ln(self.norm_a2),
            norm_t1=E.prep_ln(self.norm_t1), norm_t2=E.prep_ln(self.norm_t2),
            # operands for consumers of a LayerNorm that is never materialised (engine.LazyLN)
            cross_a_f=E.prep_folded(*E.cross_pair_weights(self.attn_a2t, self.attn_t2a), self.self_norm_a),
            cross_t_f=E.prep_folded(*E.cross_pair_weights(self.attn_t2a, self.attn_a2t), self.self_norm_t),
            ffn_a1_f=E.prep_folded(self.ffn_a[0].weight, self.ffn_a[0].bias, self.norm_a1),
            ffn_t1_f=E.prep_folded(self.ffn_t[0].weight, self.ffn_t[0].bias, self.norm_t1),
        )

    def run(self, a: E.Seq, t: E.Seq, mask_a, mask_t, want_attn: bool = False, want_f32: bool = False,
            fold_in: dict | None = None):
        """Kernel schedule of one layer on bf16 streams.  All six LayerNorms are lazy (engine.LazyLN):
        the returned streams carry norm_a2 / norm_t2 unapplied unless want_f32 asks for the tensors.
        fold_in = {"self_a": ..., "self_t": ...}: folded in-projections for lazy INPUT streams
        (built by CrossModalTransformer from the previous layer's norm_a2 / norm_t2)."""
        P = self._prep.get()
        H = self.n_heads
        maps = {}
        fold_in = fold_in or {}
        a_s, maps["audio_self"] = E.self_attention_block(a, P["self_a"], P["self_norm_a"], mask_a, H, want_attn,
                                                         lazy=True, P_folded=fold_in.get("self_a"))
        t_s, maps["text_self"] = E.self_attention_block(t, P["self_t"], P["self_norm_t"], mask_t, H, want_attn,
                                                        lazy=True, P_folded=fold_in.get("self_t"))
        qa, ka, va = E.cross_projection(a_s, P["cross_a"], P["cross_a_f"])  # a2t query | t2a key | t2a value
        qt, kt, vt_ = E.cross_projection(t_s, P["cross_t"], P["cross_t_f"])  # t2a query | a2t key | a2t value
        a1, maps["audio_queries_text"] = E.cross_attention_block(
            a_s, qa, kt, vt_, t_s.T, mask_t, P["a2t_o"]["w"], P["a2t_o"]["b"], P["norm_a1"], H, want_attn, lazy=True)
        a_o = E.ffn_block(a1, P["ffn_a1"], P["ffn_a2"], P["norm_a2"], want_f32, lazy=True, P1_folded=P["ffn_a1_f"])
        t1, maps["text_queries_audio"] = E.cross_attention_block(
            t_s, qt, ka, va, a_s.T, mask_a, P["t2a_o"]["w"], P["t2a_o"]["b"], P["norm_t1"], H, want_attn, lazy=True)
        t_o = E.ffn_block(t1, P["ffn_t1"], P["ffn_t2"], P["norm_t2"], want_f32, lazy=True, P1_folded=P["ffn_t1_f"])
        return a_o, t_o, (maps if want_attn else None)

    @torch.no_grad()
    def forward(self, h_a, h_t, mask_a=None, mask_t=None, return_attention: bool = False):
        E.warn_if_training(self, self.p_drop)
        a, t = E.to_seq(h_a, "h_a"), E.to_seq(h_t, "h_t")
        mask_a = E.check_mask(mask_a, a.B, a.T, "mask_a")
        mask_t = E.check_mask(mask_t, t.B, t.T, "mask_t")
        a_o, t_o, maps = self.run(a, t, mask_a, mask_t, return_attention, want_f32=True)
        if return_attention:
            return a_o.as_f32_3d(), t_o.as_f32_3d(), maps
        return a_o.as_f32_3d(), t_o.as_f32_3d()


class CrossModalTransformer(nn.Module):
    """Stack of CrossModalBlocks (reference :130-166)."""

    def __init__(self, num_layers=2, d_model=768, n_heads=8, dropout=0.1):
        super().__init__()
        self.layers = nn.ModuleList([CrossModalBlock(d_model, n_heads, dropout) for _ in range(num_layers)])
        self._prep = E.Prepared(self, self._build)

    def _build(self) -> list:
        """Layer i > 0 reads the previous layer's outputs through their unapplied norm_a2 / norm_t2:
        its self-attention in-projections are folded with those LayerNorms."""
        folds = [{}]
        for prev, cur in zip(self.layers[:-1], self.layers[1:]):
            folds.append(dict(
                self_a=E.prep_folded(cur.self_attn_a.in_proj_weight, cur.self_attn_a.in_proj_bias, prev.norm_a2),
                self_t=E.prep_folded(cur.self_attn_t.in_proj_weight, cur.self_attn_t.in_proj_bias, prev.norm_t2)))
        return folds

    def run(self, a: E.Seq, t: E.Seq, mask_a, mask_t, want_attn: bool = False, want_f32: bool = False):
        all_maps = []
        n = len(self.layers)
        folds = self._prep.get() if n else []
        for i, layer in enumerate(self.layers):
            a, t, maps = layer.run(a, t, mask_a, mask_t, want_attn, want_f32 and i == n - 1, fold_in=folds[i])
            if want_attn:
                all_maps.append(maps)
        return a, t, (all_maps if want_attn else None)

    @torch.no_grad()
    def forward(self, h_a, h_t, mask_a=None, mask_t=None, return_attention: bool = False):
        if len(self.layers):
            E.warn_if_training(self, self.layers[0].p_drop)
        a, t = E.to_seq(h_a, "h_a"), E.to_seq(h_t, "h_t")
        mask_a = E.check_mask(mask_a, a.B, a.T, "mask_a")
        mask_t = E.check_mask(mask_t, t.B, t.T, "mask_t")
        if not len(self.layers):
            return (h_a, h_t, []) if return_attention else (h_a, h_t)
        a, t, maps = self.run(a, t, mask_a, mask_t, return_attention, want_f32=True)
        if return_attention:
            return a.as_f32_3d(), t.as_f32_3d(), maps
        return a.as_f32_3d(), t.as_f32_3d()
