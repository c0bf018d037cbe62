"""Emotion-query Transformer decoder — drop-in for the reference's models/emotion_decoder.py
(ExplainableDecoderLayer :5-64, EmotionDecoder :66-162)."""
from __future__ import annotations

import torch
import torch.nn as nn

from hriemo import engine as E
from hriemo import lib as L
from hriemo import ops

from ._containers import MHAParams


class ExplainableDecoderLayer(nn.Module):
    """Self-attention over the N_e queries, cross-attention to the fused memory (weights
    optionally returned), FFN; post-LayerNorm x3 (reference :33-64)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1):
        super().__init__()
        self.nhead = nhead
        self.p_drop = dropout
        self.self_attn = MHAParams(d_model, nhead, dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.cross_attn = MHAParams(d_model, nhead, dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout2 = nn.Dropout(dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.activation = nn.ReLU()
        self._prep = E.Prepared(self, self._build)

    def _build(self) -> dict:
        d = self.cross_attn.embed_dim
        cw, cb = self.cross_attn.in_proj_weight.detach(), self.cross_attn.in_proj_bias.detach()
        return dict(
            self=E.prep_mha_self(self.self_attn), norm1=E.prep_ln(self.norm1),
            cross_wq=E.w16(cw[:d]), cross_bq=E.v32(cb[:d]),
            cross_wkv=E.w16(cw[d:]), cross_bkv=E.v32(cb[d:]),
            cross_wo=E.w16(self.cross_attn.out_proj.weight), cross_bo=E.v32(self.cross_attn.out_proj.bias),
            norm2=E.prep_ln(self.norm2), lin1=E.prep_linear(self.linear1), lin2=E.prep_linear(self.linear2),
            norm3=E.prep_ln(self.norm3),
        )

    def project_memory(self, mem: E.Seq) -> torch.Tensor:
        P = self._prep.get()
        return ops.gemm(mem.x, P["cross_wkv"], P["cross_bkv"], L.EPI_BIAS)  # [B*L, 2d] = [K|V]

    def run(self, zb, z32, kv_mem, mem_mask, B: int, Ne: int, Lm: int, want_attn: bool, tape: dict | None = None,
            pre_self=None, drop=None, site0: int = 0):
        return E.decoder_layer(zb, z32, kv_mem, mem_mask, self._prep.get(), B, Ne, Lm, self.nhead, want_attn, tape,
                               pre_self, drop, site0)

    @torch.no_grad()
    def forward(self, tgt, memory, memory_key_padding_mask=None, return_attention=False):
        E.warn_if_training(self, self.p_drop)
        E.require_cuda(tgt, "tgt")
        B, Ne, d = tgt.shape
        mem = E.to_seq(memory, "memory")
        mask = E.check_mask(memory_key_padding_mask, mem.B, mem.T, "memory_key_padding_mask")
        z32 = tgt.float().contiguous().view(B * Ne, d)
        zb = ops.cast_bf16(z32)
        zb, z32, probs = self.run(zb, z32, self.project_memory(mem), mask, B, Ne, mem.T, return_attention)
        return z32.view(B, Ne, d), (probs if return_attention else None)


class EmotionDecoder(nn.Module):
    """N_e learnable emotion queries decode the fused sequence; one shared Linear(d,1) head
    gives a logit per emotion (reference :117-162)."""

    def __init__(self, d_model: int = 768, num_emotions: int = 4, n_heads: int = 8, num_layers: int = 2,
                 dim_feedforward: int = 2048, dropout: float = 0.1, use_output_layer: bool = True):
        super().__init__()
        self.d_model = d_model
        self.num_emotions = num_emotions
        self.use_output_layer = use_output_layer
        self.p_drop = dropout
        self.emotion_queries = nn.Parameter(torch.randn(num_emotions, d_model))
        self.layers = nn.ModuleList([
            ExplainableDecoderLayer(d_model, n_heads, dim_feedforward, dropout) for _ in range(num_layers)])
        self.out_proj = nn.Linear(d_model, 1) if use_output_layer else None
        self._prep = E.Prepared(self, self._build)

    def _build(self) -> dict:
        p = dict(q32=E.v32(self.emotion_queries))
        if len(self.layers):
            # layer 0's self-attention block sees only the learned queries (reference :127 broadcasts the Parameter), so
            # LN(q + MHA(q, q, q)) is the same [N_e, d] matrix for every utterance: computed here once per weight set
            q32 = p["q32"]
            zb1, z32_1, _ = E.decoder_self_block(ops.cast_bf16(q32), q32, self.layers[0]._prep.get(), 1, self.num_emotions,
                                                 self.layers[0].nhead)
            p.update(self0_b=zb1, self0_f=z32_1)
        if self.out_proj is not None:
            p.update(w_out=E.v32(self.out_proj.weight), b_out=E.v32(self.out_proj.bias))
        return p

    def run(self, mem: E.Seq, mem_mask, want_attn: bool = False, tapes: list | None = None, drop=None):
        """tapes (training): a list that receives one dict of saved activations per layer (hriemo/backward.py);
        drop (training): hriemo.dropout.Drop of this step, layer i uses the sites 100000 + 100 i + 1..6."""
        P = self._prep.get()
        B, Ne, d = mem.B, self.num_emotions, self.d_model
        zb = z32 = None
        if tapes is not None or not len(self.layers):
            z32 = P["q32"].unsqueeze(0).expand(B, Ne, d).contiguous().view(B * Ne, d)  # :127 (broadcast copy)
            zb = ops.cast_bf16(z32)
        attn = []
        for li, layer in enumerate(self.layers):
            tape = None
            if tapes is not None:
                tape = {}
                tapes.append(tape)
            pre_self = None
            if li == 0 and tapes is None:   # inference: the precomputed block, broadcast over the batch (training keeps
                # the full schedule: the backward pass needs the block's activations for its parameter gradients)
                pre_self = (P["self0_b"].unsqueeze(0).expand(B, Ne, d).contiguous().view(B * Ne, d),
                            P["self0_f"].unsqueeze(0).expand(B, Ne, d).contiguous().view(B * Ne, d))
            zb, z32, probs = layer.run(zb, z32, layer.project_memory(mem), mem_mask, B, Ne, mem.T, want_attn, tape,
                                       pre_self, drop, 100000 + 100 * li)
            if want_attn and probs is not None:
                attn.append(probs)
        logits = None
        if self.out_proj is not None:  # :153-155: shared head, squeeze(-1)
            logits = ops.sgemm(z32, P["w_out"], P["b_out"], L.ACT_NONE).view(B, Ne)
        return z32.view(B, Ne, d), logits, (attn if want_attn else None)

    @torch.no_grad()
    def forward(self, memory, memory_key_padding_mask=None, return_attention: bool = False):
        E.warn_if_training(self, self.p_drop)
        mem = E.to_seq(memory, "memory")
        mask = E.check_mask(memory_key_padding_mask, mem.B, mem.T, "memory_key_padding_mask")
        z, logits, attn = self.run(mem, mask, return_attention)
        if return_attention:
            return z, logits, attn
        return z, logits
