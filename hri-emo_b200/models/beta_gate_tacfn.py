"""Vector-wise beta-gate — drop-in for the reference's models/beta_gate_tacfn.py
(masked_mean :6-24, BetaGate :27-118)."""
from __future__ import annotations

import torch
import torch.nn as nn

from hriemo import engine as E
from hriemo import lib as L
from hriemo import ops


def masked_mean(x: torch.Tensor, mask: torch.Tensor | None) -> torch.Tensor:
    """Mean over time with an optional PAD mask (True = PAD); reference :6-24.
    x: [B, L, d] on a CUDA device (fp32 or bf16) -> [B, d] fp32."""
    s = E.to_seq(x, "masked_mean x")
    mask = E.check_mask(mask, s.B, s.T, "mask")
    return ops.ln_masked_mean(s.x, None, None, mask, s.B, s.T, apply_ln=False)


class BetaGate(nn.Module):
    """w = sigmoid(MLP([a, t, |a-t|, a*t])) on LayerNorm-ed, masked-mean-pooled streams;
    h = w*LN(a)[:, :L] + (1-w)*LN(t) with L = T_t; beta = mean_d(w).  Reference :68-118."""

    def __init__(self, d_model: int = 768, hidden_dim: int = 256):
        super().__init__()
        self.d_model = d_model
        self.norm_a = nn.LayerNorm(d_model)
        self.norm_t = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(nn.Linear(d_model * 4, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, d_model))
        self._prep = E.Prepared(self, self._build)

    def _build(self) -> dict:
        # the gate MLP stays fp32: beta feeds a bit-sensitive decision
        P = dict(norm_a=E.prep_ln(self.norm_a), norm_t=E.prep_ln(self.norm_t),
                 w0=E.v32(self.mlp[0].weight), b0=E.v32(self.mlp[0].bias),
                 w2=E.v32(self.mlp[2].weight), b2=E.v32(self.mlp[2].bias))
        # The first layer ([B, 4d] -> hidden) is 90 % of the MLP's arithmetic: it runs on the tcgen05 GEMM with both
        # operands split into bf16 hi + lo (ops.split3: 16 mantissa bits, fp32 accumulation -- hriemo/precise.py), which
        # keeps beta within 1e-7 of the fp32 FMA loop (0.24 ms -> ~0.03 ms per 2048 utterances).  Needs hidden % 32 == 0.
        if P["w0"].shape[0] % 32 == 0:
            P["w0_3"] = ops.split3(P["w0"], weight=True)
        return P

    def run(self, a: E.Seq, t: E.Seq, mask_a, mask_t, want_bf16: bool = True, want_f32: bool = False):
        if a.T != t.T and a.T < t.T:
            # reference :107-116 slices audio to L = T_t and then fails to broadcast
            raise RuntimeError(f"BetaGate: audio length {a.T} is shorter than text length {t.T}")
        P = self._prep.get()
        # streams that arrive with an unapplied encoder LayerNorm (engine.LazyLN) get it applied per
        # row inside the gate kernels, in front of norm_a / norm_t
        pre_a = None if a.ln is None else (a.ln.gamma, a.ln.beta, a.ln.stats)
        pre_t = None if t.ln is None else (t.ln.gamma, t.ln.beta, t.ln.stats)
        a_pool = ops.ln_masked_mean(a.x, *P["norm_a"], mask_a, a.B, a.T, pre_ln=pre_a)  # :79, :83
        t_pool = ops.ln_masked_mean(t.x, *P["norm_t"], mask_t, t.B, t.T, pre_ln=pre_t)  # :80, :84
        g = ops.gate_input(a_pool, t_pool)                                 # :87-89
        if "w0_3" in P:
            hid = ops.gemm(ops.split3(g), P["w0_3"], P["b0"], L.EPI_BIAS_F32)   # pre-activation; ReLU on the way into layer 2
            w = ops.sgemm(hid, P["w2"], P["b2"], L.ACT_SIGMOID | L.ACT_RELU_IN)  # :92
        else:
            hid = ops.sgemm(g, P["w0"], P["b0"], L.ACT_RELU)
            w = ops.sgemm(hid, P["w2"], P["b2"], L.ACT_SIGMOID)                # :92
        hb, hf, beta = ops.gate_blend(a.x, a.T, t.x, P["norm_a"], P["norm_t"], w, a.B, t.T,
                                      want_bf16=want_bf16, want_f32=want_f32, pre_ln_a=pre_a,
                                      pre_ln_t=pre_t)  # :95-116
        return E.Seq(hb, a.B, t.T, hf), beta

    @torch.no_grad()
    def forward(self, h_a, h_t, mask_a=None, mask_t=None):
        a, t = E.to_seq(h_a, "h_a"), E.to_seq(h_t, "h_t")
        mask_a = E.check_mask(mask_a, a.B, a.T, "mask_a")
        mask_t = E.check_mask(mask_t, t.B, t.T, "mask_t")
        h, beta = self.run(a, t, mask_a, mask_t, want_bf16=False, want_f32=True)
        return h.as_f32_3d(), beta
