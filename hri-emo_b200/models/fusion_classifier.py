"""Baseline classifier head — drop-in for the reference's models/fusion_classifier.py (:10-150):
TACFN encoder + vector beta-gate + UNMASKED mean over time + LN/Linear/ReLU/Linear."""
from __future__ import annotations

import torch
import torch.nn as nn

from hriemo import engine as E
from hriemo import lib as L
from hriemo import ops

from .beta_gate_tacfn import BetaGate
from .cross_modal_block_tacfn import CrossModalTransformer


class FusionClassifier(nn.Module):
    def __init__(self, d_model: int = 768, num_classes: int = 4, n_heads: int = 8, num_layers: int = 2,
                 beta_hidden: int = 256, dropout: float = 0.2):
        super().__init__()
        self.p_drop = dropout
        self.cross_modal = CrossModalTransformer(num_layers=num_layers, d_model=d_model, n_heads=n_heads,
                                                 dropout=dropout)
        self.beta_gate = BetaGate(d_model=d_model, hidden_dim=beta_hidden)
        self.classifier = nn.Sequential(nn.LayerNorm(d_model), nn.Linear(d_model, d_model), nn.ReLU(),
                                        nn.Dropout(dropout), nn.Linear(d_model, num_classes))
        self._prep = E.Prepared(self, self._build)

    def _build(self) -> dict:
        c = self.classifier
        return dict(ln=E.prep_ln(c[0]), w1=E.v32(c[1].weight), b1=E.v32(c[1].bias),
                    w2=E.v32(c[4].weight), b2=E.v32(c[4].bias))

    def _ensure_3d(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 2:
            return x.unsqueeze(1)
        elif x.dim() == 3:
            return x
        else:
            raise ValueError(f"Expected 2D or 3D tensor, got shape {x.shape}")

    @torch.no_grad()
    def forward(self, h_a, h_t, mask_a=None, mask_t=None):
        E.warn_if_training(self, self.p_drop)
        P = self._prep.get()
        a = E.to_seq(self._ensure_3d(h_a), "h_a")
        t = E.to_seq(self._ensure_3d(h_t), "h_t")
        mask_a = E.check_mask(mask_a, a.B, a.T, "mask_a")
        mask_t = E.check_mask(mask_t, t.B, t.T, "mask_t")
        a, t, _ = self.cross_modal.run(a, t, mask_a, mask_t)
        h, beta = self.beta_gate.run(a, t, mask_a, mask_t, want_bf16=False, want_f32=True)
        pooled = ops.mean_over_time(h.x32, h.B, h.T)                    # reference :145 (unmasked)
        _, x = ops.layernorm(pooled, *P["ln"], want_bf16=False, want_f32=True)
        x = ops.sgemm(x, P["w1"], P["b1"], L.ACT_RELU)
        logits = ops.sgemm(x, P["w2"], P["b2"], L.ACT_NONE)
        return logits, beta, pooled
