"""MOSEI wrapper — drop-in for the reference's models/mosei_fusion_with_emotion_decoder.py
(:7-79): Linear projections of COVAREP / GloVe features to d_model, then the backbone."""
from __future__ import annotations

import torch
import torch.nn as nn

from hriemo import engine as E
from hriemo import lib as L
from hriemo import ops
from hriemo import precise

from .fusion_with_emotion_decoder import FusionWithEmotionDecoder


class MoseiFusionWithEmotionDecoder(nn.Module):
    def __init__(self, d_audio: int, d_text: int, d_model: int = 256, num_emotions: int = 6, n_heads: int = 4,
                 num_layers_fusion: int = 2, num_layers_decoder: int = 2, beta_hidden: int = 128,
                 dropout: float = 0.2):
        super().__init__()
        self.d_audio, self.d_text = d_audio, d_text
        self.audio_proj = nn.Linear(d_audio, d_model)
        self.text_proj = nn.Linear(d_text, d_model)
        self.backbone = FusionWithEmotionDecoder(d_model=d_model, num_emotions=num_emotions, n_heads=n_heads,
                                                 num_layers_fusion=num_layers_fusion,
                                                 num_layers_decoder=num_layers_decoder,
                                                 beta_hidden=beta_hidden, dropout=dropout)
        self._prep = E.Prepared(self, self._build)

    def _build(self) -> dict:
        # K (= 74 / 300) is zero-padded to a multiple of 8 so rows are 16-byte aligned for TMA
        ka, kt = ops.round_up(self.d_audio, 8), ops.round_up(self.d_text, 8)
        return dict(ka=ka, kt=kt, a=E.prep_linear(self.audio_proj, ka), t=E.prep_linear(self.text_proj, kt))

    @torch.no_grad()
    def project_inputs(self, h_a, h_t, keep_inputs: bool = False):
        """audio_proj / text_proj (reference :64-66) over the whole batch -> bf16 [B, T, d_model] tensors (and, with
        keep_inputs, the zero-padded bf16 inputs [B*T, K8] the weight gradients are formed from)."""
        P = self._prep.get()
        # training: the kept inputs are zero-padded to a multiple of 128 columns (what the tcgen05 wgrad kernel tiles);
        # the forward GEMM reads their first K8 columns as a strided view
        la = ops.round_up(self.d_audio, 128) if keep_inputs else P["ka"]
        lt = ops.round_up(self.d_text, 128) if keep_inputs else P["kt"]
        a, t = E.to_seq(h_a, "h_a", la), E.to_seq(h_t, "h_t", lt)
        pa = ops.gemm(a.x[:, :P["ka"]], P["a"]["w"], P["a"]["b"], L.EPI_BIAS)
        pt = ops.gemm(t.x[:, :P["kt"]], P["t"]["w"], P["t"]["b"], L.EPI_BIAS)
        xa, xt = pa.view(a.B, a.T, -1), pt.view(t.B, t.T, -1)
        return (xa, xt, a.x, t.x) if keep_inputs else (xa, xt)

    def forward(self, h_a, h_t, mask_a=None, mask_t=None, return_attention=False):
        """Same contract as the backbone's forward; in train() mode with gradients enabled the outputs carry a grad_fn
        whose backward also fills audio_proj / text_proj (hriemo/autograd.py), so the MOSEI training loop
        (scripts/fusion/train_mosei_fusion_seq_level_decoder.py:367-429: autocast, pos_weight BCE, beta entropy,
        GradScaler, gradient accumulation) runs unmodified."""
        E.require_cuda(h_a, "h_a")
        E.require_cuda(h_t, "h_t")
        mask_a = E.check_mask(mask_a, h_a.shape[0], h_a.shape[1], "mask_a")
        mask_t = E.check_mask(mask_t, h_t.shape[0], h_t.shape[1], "mask_t")
        if precise.get_mode() == "tf32x3":   # hriemo/precise.py; inference only
            logits, beta, z, pack = precise.mosei_forward(self, h_a, h_t, mask_a, mask_t, return_attention)
            return (logits, beta, z, pack) if return_attention else (logits, beta, z)
        if (self.training and torch.is_grad_enabled() and not return_attention
                and any(p.requires_grad for p in self.parameters())):
            from hriemo.autograd import mosei_forward_with_grad
            return mosei_forward_with_grad(self, h_a, h_t, mask_a, mask_t)
        E.warn_if_training(self, self.backbone.p_drop)   # the inference schedule: no dropout even in train() mode
        with torch.no_grad():
            return self._forward_eval(h_a, h_t, mask_a, mask_t, return_attention)

    def _forward_eval(self, h_a, h_t, mask_a, mask_t, return_attention):
        P = self._prep.get()

        def project(a: E.Seq, t: E.Seq):  # reference :64-66
            pa = ops.gemm(a.x, P["a"]["w"], P["a"]["b"], L.EPI_BIAS)
            pt = ops.gemm(t.x, P["t"]["w"], P["t"]["b"], L.EPI_BIAS)
            return E.Seq(pa, a.B, a.T), E.Seq(pt, t.B, t.T)

        logits, beta, z, pack = self.backbone.run_slabbed(h_a, h_t, mask_a, mask_t, return_attention,
                                                          ld_a=P["ka"], ld_t=P["kt"], pre=project)
        if return_attention:
            return logits, beta, z, pack
        return logits, beta, z
