"""Stock-PyTorch competitor arm: the fusion-and-decode forward composed from torch.nn modules only.

BENCH INFRASTRUCTURE.  The reference (Makiato1999/HRI-EMO) ships no kernel of its own: on a GPU it runs
`nn.MultiheadAttention` (cuBLASLt in/out projections + SDPA, or `_native_multi_head_attention` on the
eval()/no_grad self-attention fast path), `nn.LayerNorm`, `nn.Linear`, `torch.sigmoid` -- i.e. stock
PyTorch eager.  The reference package itself cannot travel to the GPU box, so this file rebuilds
the SAME module graph from a `state_dict` (whatever the key names say is there: every `*.in_proj_weight`
becomes an `nn.MultiheadAttention(batch_first=True)`, every 1-D `weight` an `nn.LayerNorm`, every 2-D one an
`nn.Linear`) and drives it in the reference's order of calls, with the reference's `need_weights=False`
(models/cross_modal_block_tacfn.py:74-118, models/emotion_decoder.py:42-59), so the same ATen kernels are
dispatched.  tests/test_stock_torch_cpu.py checks it against the golden outputs of the unmodified reference.

It is timed by `bench.py --impl torch_gpu` and the `gpu_eager_baseline` key; nothing in `hri-emo_b200/`
imports it, and it calls none of this repository's kernels.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn


def _modules_from_state(sd: Dict[str, torch.Tensor], n_heads: int) -> nn.ModuleDict:
    mods = {}
    for key, w in sd.items():
        if key.endswith(".in_proj_weight"):
            p = key[: -len(".in_proj_weight")]
            m = nn.MultiheadAttention(w.shape[1], n_heads, dropout=0.0, batch_first=True)
            m.load_state_dict({k[len(p) + 1:]: v for k, v in sd.items() if k.startswith(p + ".")})
            mods[p] = m
    for key, w in sd.items():
        if not key.endswith(".weight"):
            continue
        p = key[: -len(".weight")]
        if any(p == a + ".out_proj" for a in mods):      # MultiheadAttention's own out_proj
            continue
        b = sd.get(p + ".bias")
        if w.dim() == 1:
            m = nn.LayerNorm(w.shape[0])
        else:
            m = nn.Linear(w.shape[1], w.shape[0], bias=b is not None)
        m.load_state_dict({"weight": w} if b is None else {"weight": w, "bias": b})
        mods[p] = m
    return nn.ModuleDict({k.replace(".", "/"): v for k, v in mods.items()})


class StockFusion(nn.Module):
    """FusionWithEmotionDecoder / MoseiFusionWithEmotionDecoder forward on stock torch.nn modules."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], n_heads: int):
        super().__init__()
        sd = {k: v.detach().clone() for k, v in state_dict.items()}
        self.wrapped = any(k.startswith("backbone.") for k in sd)
        self.m = _modules_from_state(sd, n_heads)
        root = "backbone." if self.wrapped else ""
        self.root = root
        self.queries = nn.Parameter(sd[root + "emotion_decoder.emotion_queries"])
        self.n_fusion = 1 + max(int(k.split(".")[len(root.split(".")) + 1]) for k in sd if k.startswith(root + "cross_modal.layers."))
        self.n_decoder = 1 + max(int(k.split(".")[len(root.split(".")) + 1]) for k in sd if k.startswith(root + "emotion_decoder.layers."))

    def mod(self, name: str) -> nn.Module:
        return self.m[(self.root + name).replace(".", "/")]

    def attend(self, name: str, q, kv, pad):
        return self.mod(name)(q, kv, kv, key_padding_mask=pad, need_weights=False)[0]

    def ffn(self, name: str, x):
        return self.mod(name + ".2")(torch.relu(self.mod(name + ".0")(x)))

    def forward(self, h_a, h_t, mask_a: Optional[torch.Tensor] = None, mask_t: Optional[torch.Tensor] = None):
        if self.wrapped:
            h_a = self.m["audio_proj"](h_a)
            h_t = self.m["text_proj"](h_t)
        a, t = h_a, h_t
        for i in range(self.n_fusion):
            L = f"cross_modal.layers.{i}."
            a = self.mod(L + "self_norm_a")(a + self.attend(L + "self_attn_a", a, a, mask_a))
            t = self.mod(L + "self_norm_t")(t + self.attend(L + "self_attn_t", t, t, mask_t))
            a1 = self.mod(L + "norm_a1")(a + self.attend(L + "attn_a2t", a, t, mask_t))
            t1 = self.mod(L + "norm_t1")(t + self.attend(L + "attn_t2a", t, a, mask_a))
            a = self.mod(L + "norm_a2")(a1 + self.ffn(L + "ffn_a", a1))
            t = self.mod(L + "norm_t2")(t1 + self.ffn(L + "ffn_t", t1))
        # vector gate
        a_n, t_n = self.mod("beta_gate.norm_a")(a), self.mod("beta_gate.norm_t")(t)

        def pooled(x, pad):
            if pad is None:
                return x.mean(dim=1)
            keep = (~pad).unsqueeze(-1).to(x.dtype)
            return (x * keep).sum(dim=1) / keep.sum(dim=1).clamp(min=1.0)

        pa, pt = pooled(a_n, mask_a), pooled(t_n, mask_t)
        g = torch.cat([pa, pt, (pa - pt).abs(), pa * pt], dim=-1)
        w = torch.sigmoid(self.mod("beta_gate.mlp.2")(torch.relu(self.mod("beta_gate.mlp.0")(g))))
        beta = w.mean(dim=-1, keepdim=True)
        Lf = t_n.shape[1]
        wb = w.unsqueeze(1)
        fused = wb * a_n[:, :Lf] + (1.0 - wb) * t_n
        fmask = None
        if mask_a is not None or mask_t is not None:
            B = fused.shape[0]
            fmask = torch.zeros((B, Lf), dtype=torch.bool, device=fused.device)
            if mask_a is not None:
                ma = mask_a[:, :Lf]
                if ma.shape[1] < Lf:
                    ma = torch.cat([ma, torch.ones((B, Lf - ma.shape[1]), dtype=torch.bool, device=ma.device)], dim=1)
                fmask = fmask | ma
            if mask_t is not None:
                fmask = fmask | mask_t
        z = self.queries.unsqueeze(0).expand(fused.shape[0], -1, -1)
        for i in range(self.n_decoder):
            L = f"emotion_decoder.layers.{i}."
            z = self.mod(L + "norm1")(z + self.attend(L + "self_attn", z, z, None))
            z = self.mod(L + "norm2")(z + self.attend(L + "cross_attn", z, fused, fmask))
            z = self.mod(L + "norm3")(z + self.mod(L + "linear2")(torch.relu(self.mod(L + "linear1")(z))))
        logits = self.mod("emotion_decoder.out_proj")(z).squeeze(-1)
        return logits, beta, z


MODES = ("fp32", "tf32", "bf16_autocast")


def run_mode(model: StockFusion, mode: str, h_a, h_t, mask_a=None, mask_t=None):
    """One forward in one of the three precisions the reference can run on a GPU:
    fp32 with TF32 off, TF32 on, torch.autocast(bfloat16) (scripts/infer/mosei_eval_infer.py:195-197)."""
    tf32 = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad():
        if mode == "bf16_autocast":
            with torch.autocast(device_type=h_a.device.type, dtype=torch.bfloat16):
                return model(h_a, h_t, mask_a, mask_t)
        return model(h_a, h_t, mask_a, mask_t)
