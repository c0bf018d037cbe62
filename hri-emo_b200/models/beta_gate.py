"""Legacy scalar beta-gate — drop-in for the reference's models/beta_gate.py
(masked_mean :6-33, BetaGate :36-114)."""
from __future__ import annotations

import torch
import torch.nn as nn

from hriemo import engine as E
from hriemo import lib as L
from hriemo import ops

from .beta_gate_tacfn import masked_mean  # same semantics as reference :6-33

__all__ = ["masked_mean", "BetaGate"]


class BetaGate(nn.Module):
    """beta = sigmoid(MLP([a, t, |a-t|, a*t])) in [B,1]; h = beta*a[:, :L] + (1-beta)*t (no LayerNorm)."""

    def __init__(self, d_model=768, hidden_dim=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(d_model * 4, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, 1))
        self._prep = E.Prepared(self, self._build)

    def _build(self) -> dict:
        return dict(w0=E.v32(self.mlp[0].weight), b0=E.v32(self.mlp[0].bias),
                    w2=E.v32(self.mlp[2].weight), b2=E.v32(self.mlp[2].bias))

    def run(self, a: E.Seq, t: E.Seq, mask_a, mask_t, want_bf16: bool = True, want_f32: bool = False):
        if a.T != t.T and a.T < t.T:
            raise RuntimeError(f"BetaGate: audio length {a.T} is shorter than text length {t.T}")
        P = self._prep.get()
        pre_a = None if a.ln is None else (a.ln.gamma, a.ln.beta, a.ln.stats)  # unapplied encoder LayerNorm (engine.LazyLN)
        pre_t = None if t.ln is None else (t.ln.gamma, t.ln.beta, t.ln.stats)
        a_pool = ops.ln_masked_mean(a.x, None, None, mask_a, a.B, a.T, apply_ln=False, pre_ln=pre_a)  # :81
        t_pool = ops.ln_masked_mean(t.x, None, None, mask_t, t.B, t.T, apply_ln=False, pre_ln=pre_t)  # :82
        g = ops.gate_input(a_pool, t_pool)                                               # :85-87
        hid = ops.sgemm(g, P["w0"], P["b0"], L.ACT_RELU)
        beta = ops.sgemm(hid, P["w2"], P["b2"], L.ACT_SIGMOID)                           # :90  [B,1]
        hb, hf, beta_out = ops.gate_blend(a.x, a.T, t.x, None, None, beta, a.B, t.T, apply_ln=False,
                                          w_is_scalar=True, want_bf16=want_bf16, want_f32=want_f32,
                                          pre_ln_a=pre_a, pre_ln_t=pre_t)
        return E.Seq(hb, a.B, t.T, hf), beta_out

    @torch.no_grad()
    def forward(self, h_a, h_t, mask_a=None, mask_t=None):
        a, t = E.to_seq(h_a, "h_a"), E.to_seq(h_t, "h_t")
        mask_a = E.check_mask(mask_a, a.B, a.T, "mask_a")
        mask_t = E.check_mask(mask_t, t.B, t.T, "mask_t")
        h, beta = self.run(a, t, mask_a, mask_t, want_bf16=False, want_f32=True)
        return h.as_f32_3d(), beta
