"""Full seq-level model — drop-in for the reference's models/fusion_with_emotion_decoder.py
(FusionWithEmotionDecoder :10-197): cross-modal encoder -> vector beta-gate -> emotion decoder."""
from __future__ import annotations

import torch
import torch.nn as nn

from hriemo import engine as E
from hriemo import precise

from .beta_gate_tacfn import BetaGate
from .cross_modal_block_tacfn import CrossModalTransformer
from .emotion_decoder import EmotionDecoder

# Utterances are processed in slabs so that the widest intermediate (the 4d FFN hidden,
# bf16) stays bounded whatever the batch size; every kernel still sees >= 148 tiles.
# HRIEMO_MAX_ROWS_PER_SLAB overrides the row budget (an A / B switch of the benchmark).
import os as _os

MAX_ROWS_PER_SLAB = int(_os.environ.get("HRIEMO_MAX_ROWS_PER_SLAB", 1 << 20))


class FusionWithEmotionDecoder(nn.Module):
    def __init__(self, d_model: int = 768, num_emotions: int = 4, n_heads: int = 8,
                 num_layers_fusion: int = 2, num_layers_decoder: int = 2, beta_hidden: int = 256,
                 dropout: float = 0.1):
        super().__init__()
        self.p_drop = dropout
        self.cross_modal = CrossModalTransformer(num_layers=num_layers_fusion, d_model=d_model,
                                                 n_heads=n_heads, dropout=dropout)
        self.beta_gate = BetaGate(d_model=d_model, hidden_dim=beta_hidden)
        self.emotion_decoder = EmotionDecoder(d_model=d_model, num_emotions=num_emotions, n_heads=n_heads,
                                              num_layers=num_layers_decoder, dropout=dropout,
                                              use_output_layer=True)

    # ----------------------------------------------------------------- helpers
    def _ensure_3d(self, x: torch.Tensor) -> torch.Tensor:
        """[B,d] -> [B,1,d] (reference :60-69)."""
        if x.dim() == 2:
            return x.unsqueeze(1)
        if x.dim() == 3:
            return x
        raise ValueError(f"Expected 2D or 3D tensor, got {x.shape}")

    def _build_fused_mask(self, mask_a, mask_t, L_fused: int):
        """PAD where either modality is PAD, aligned to the fused length (reference :71-115).
        Boolean index bookkeeping on [B, L] masks: host-side torch ops, not a kernel."""
        if mask_a is None and mask_t is None:
            return None

        def fit(m):
            if m is None:
                return None
            if m.size(1) < L_fused:
                pad = torch.ones(m.size(0), L_fused - m.size(1), dtype=torch.bool, device=m.device)
                return torch.cat([m, pad], dim=1)
            return m[:, :L_fused]

        ma, mt = fit(mask_a), fit(mask_t)
        if ma is None:
            return mt.contiguous()
        if mt is None:
            return ma.contiguous()
        return (ma | mt).contiguous()

    # ----------------------------------------------------------------- kernel schedule
    def run(self, a: E.Seq, t: E.Seq, mask_a, mask_t, want_attn: bool = False):
        a, t, enc_attn = self.cross_modal.run(a, t, mask_a, mask_t, want_attn)          # :145-156
        h, beta = self.beta_gate.run(a, t, mask_a, mask_t)                               # :159
        fused_mask = self._build_fused_mask(mask_a, mask_t, h.T)                         # :165
        z, logits, dec_attn = self.emotion_decoder.run(h, fused_mask, want_attn)         # :170-184
        return logits, beta, z, ({"encoder": enc_attn, "decoder": dec_attn} if want_attn else None)

    def run_slabbed(self, h_a, h_t, mask_a, mask_t, want_attn: bool, ld_a=None, ld_t=None, pre=None):
        """Split the batch into slabs of utterances (independent: SURVEY sec. 8e) and run each."""
        B, T_a = h_a.shape[0], h_a.shape[1]
        per = max(1, MAX_ROWS_PER_SLAB // max(T_a, 1))
        outs = []
        for s in range(0, B, per):
            e = min(B, s + per)
            a = E.to_seq(h_a[s:e], "h_a", ld_a)
            t = E.to_seq(h_t[s:e], "h_t", ld_t)
            if pre is not None:
                a, t = pre(a, t)
            ma = None if mask_a is None else mask_a[s:e].contiguous()
            mt = None if mask_t is None else mask_t[s:e].contiguous()
            outs.append(self.run(a, t, ma, mt, want_attn))
        if len(outs) == 1:
            return outs[0]
        logits = torch.cat([o[0] for o in outs], dim=0)
        beta = torch.cat([o[1] for o in outs], dim=0)
        z = torch.cat([o[2] for o in outs], dim=0)
        pack = None
        if want_attn:
            n_enc = len(outs[0][3]["encoder"])
            enc = [{k: torch.cat([o[3]["encoder"][i][k] for o in outs], dim=0) for k in outs[0][3]["encoder"][i]}
                   for i in range(n_enc)]
            dec = [torch.cat([o[3]["decoder"][i] for o in outs], dim=0) for i in range(len(outs[0][3]["decoder"]))]
            pack = {"encoder": enc, "decoder": dec}
        return logits, beta, z, pack

    def forward(self, h_a, h_t, mask_a=None, mask_t=None, return_attention: bool = False):
        """h_a [B,d] or [B,L_a,d], h_t [B,d] or [B,L_t,d]; masks bool, True = PAD.
        Returns (logits [B,N_e], beta [B,1], z [B,N_e,d]) or, with return_attention, a 4-tuple
        whose last item is {"encoder": [...per layer dict...], "decoder": [...per layer...]}.

        In train() mode with gradients enabled the three outputs carry a grad_fn (hriemo/autograd.py): the reference's
        training loops -- criterion(logits, y) [+ beta terms], loss.backward(), clip_grad_norm_, optimizer.step()
        (scripts/fusion/train_fusion_seq_level_decoder.py:310-333) -- run unmodified and fill .grad through the
        hand-scheduled backward pass; dropout (p > 0) is applied at every site of the reference, hriemo/dropout.py)."""
        h_a = self._ensure_3d(h_a)
        h_t = self._ensure_3d(h_t)
        E.require_cuda(h_a, "h_a")
        E.require_cuda(h_t, "h_t")
        mask_a = E.check_mask(mask_a, h_a.shape[0], h_a.shape[1], "mask_a")
        mask_t = E.check_mask(mask_t, h_t.shape[0], h_t.shape[1], "mask_t")
        if precise.get_mode() == "tf32x3":   # fp32 activations, split-bf16 GEMMs (hriemo/precise.py); inference only
            logits, beta, z, pack = precise.fusion_forward(self, h_a, h_t, mask_a, mask_t, return_attention)
            return (logits, beta, z, pack) if return_attention else (logits, beta, z)
        if (self.training and torch.is_grad_enabled() and not return_attention
                and any(p.requires_grad for p in self.parameters())):
            from hriemo.autograd import fusion_forward_with_grad
            return fusion_forward_with_grad(self, h_a, h_t, mask_a, mask_t)
        E.warn_if_training(self, self.p_drop)   # the inference schedule: no dropout even in train() mode
        with torch.no_grad():
            logits, beta, z, pack = self.run_slabbed(h_a, h_t, mask_a, mask_t, return_attention)
        if return_attention:
            return logits, beta, z, pack
        return logits, beta, z
