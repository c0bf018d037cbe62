"""Legacy bidirectional cross-modal block (no intra-modal stage) — drop-in for the
reference's models/cross_modal_block.py (CrossModalBlock :5-71, CrossModalTransformer :74-95)."""
from __future__ import annotations

import torch
import torch.nn as nn

from hriemo import engine as E

from ._containers import MHAParams, ffn


class CrossModalBlock(nn.Module):
    def __init__(self, d_model=768, n_heads=8, dropout=0.1):
        super().__init__()
        self.n_heads = n_heads
        self.p_drop = dropout
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
            cross_a=E.prep_cross_pair(self.attn_a2t, self.attn_t2a),
            cross_t=E.prep_cross_pair(self.attn_t2a, self.attn_a2t),
            a2t_o=E.prep_linear(self.attn_a2t.out_proj), t2a_o=E.prep_linear(self.attn_t2a.out_proj),
            ffn_a1=E.prep_linear(self.ffn_a[0]), ffn_a2=E.prep_linear(self.ffn_a[2]),
            ffn_t1=E.prep_linear(self.ffn_t[0]), ffn_t2=E.prep_linear(self.ffn_t[2]),
            norm_a1=E.prep_ln(self.norm_a1), norm_a2=E.prep_ln(self.norm_a2),
            norm_t1=E.prep_ln(self.norm_t1), norm_t2=E.prep_ln(self.norm_t2),
        )

    def run(self, a: E.Seq, t: E.Seq, mask_a, mask_t, want_f32: bool = False):
        """Reference :56-69 — both directions read the layer inputs."""
        P = self._prep.get()
        H = self.n_heads
        qa, ka, va = E.cross_projection(a, P["cross_a"])
        qt, kt, vt_ = E.cross_projection(t, P["cross_t"])
        a1, _ = E.cross_attention_block(a, qa, kt, vt_, t.T, mask_t, P["a2t_o"]["w"], P["a2t_o"]["b"],
                                        P["norm_a1"], H, False)
        a_o = E.ffn_block(a1, P["ffn_a1"], P["ffn_a2"], P["norm_a2"], want_f32)
        t1, _ = E.cross_attention_block(t, qt, ka, va, a.T, mask_a, P["t2a_o"]["w"], P["t2a_o"]["b"],
                                        P["norm_t1"], H, False)
        t_o = E.ffn_block(t1, P["ffn_t1"], P["ffn_t2"], P["norm_t2"], want_f32)
        return a_o, t_o

    @torch.no_grad()
    def forward(self, h_a, h_t, mask_a=None, mask_t=None):
        E.warn_if_training(self, self.p_drop)
        a, t = E.to_seq(h_a, "h_a"), E.to_seq(h_t, "h_t")
        mask_a = E.check_mask(mask_a, a.B, a.T, "mask_a")
        mask_t = E.check_mask(mask_t, t.B, t.T, "mask_t")
        a_o, t_o = self.run(a, t, mask_a, mask_t, want_f32=True)
        return a_o.as_f32_3d(), t_o.as_f32_3d()


class CrossModalTransformer(nn.Module):
    def __init__(self, num_layers=2, d_model=768, n_heads=8, dropout=0.1):
        super().__init__()
        self.layers = nn.ModuleList([CrossModalBlock(d_model, n_heads, dropout) for _ in range(num_layers)])

    def run(self, a: E.Seq, t: E.Seq, mask_a, mask_t, want_f32: bool = False):
        n = len(self.layers)
        for i, layer in enumerate(self.layers):
            a, t = layer.run(a, t, mask_a, mask_t, want_f32 and i == n - 1)
        return a, t

    @torch.no_grad()
    def forward(self, h_a, h_t, mask_a=None, mask_t=None):
        if not len(self.layers):
            return h_a, h_t
        E.warn_if_training(self, self.layers[0].p_drop)
        a, t = E.to_seq(h_a, "h_a"), E.to_seq(h_t, "h_t")
        mask_a = E.check_mask(mask_a, a.B, a.T, "mask_a")
        mask_t = E.check_mask(mask_t, t.B, t.T, "mask_t")
        a, t = self.run(a, t, mask_a, mask_t, want_f32=True)
        return a.as_f32_3d(), t.as_f32_3d()
