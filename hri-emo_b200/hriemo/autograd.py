"""The training boundary of the drop-in modules: `model(...)` in train() mode returns tensors that carry a grad_fn, and
`loss.backward()` on anything computed from them runs the hand-scheduled backward pass (hriemo/backward.py) and fills
`.grad` of the module's parameters -- so the reference's own training loops run unmodified on the B200 path:

    logits, beta, z = model(h_a, h_t, m_a, m_t)                    scripts/fusion/train_fusion_seq_level_decoder.py:310
    loss = criterion(logits, y) (+ beta regulariser / entropy)     :313-327, train_mosei_...:385-387 (pos_weight, H(beta))
    loss.backward(); clip_grad_norm_(...); optimizer.step()        :331-333 (or through a GradScaler: train_mosei_...:395-402)

The loss, its pos_weight, the beta terms, the NaN / Inf batch skip, gradient accumulation and the optimizer are
therefore the CALLER'S torch code, exactly as in the reference; what runs on the C-ABI kernels is the model's forward
with tapes and its backward from (d_logits, d_beta, d_z).  Parameters enter the autograd.Function as inputs, which is
what makes torch accumulate into their `.grad`.

Dropout: applied at every site of the reference when the model was built with p > 0 (hriemo/dropout.py: the forward draws a
seed from torch's CPU generator, the backward recomputes the masks from the same stream keys).
hriemo.train.Trainer remains the fast path (flat arenas, fused clip / AdamW, CUDA-graph replay)."""
from __future__ import annotations

from typing import Optional

import torch

from . import backward as BW
from . import ops


class _FusionFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, names, h_a, h_t, mask_a, mask_t, *params):
        with torch.no_grad():
            logits, beta, z, tape = BW.forward_train(model, h_a, h_t, mask_a, mask_t)
        ctx.model, ctx.names, ctx.tape = model, names, tape
        return logits, beta, z

    @staticmethod
    def backward(ctx, d_logits, d_beta, d_z):
        with torch.no_grad():
            grads, _, _ = BW.backward_from(ctx.model, ctx.tape, d_logits, d_beta, d_z)
        ctx.tape = None
        out = []
        for n in ctx.names:
            g = grads.get(n)
            out.append(None if g is None else g.reshape(dict(ctx.model.named_parameters())[n].shape))
        return (None, None, None, None, None, None) + tuple(out)


def fusion_forward_with_grad(model, h_a, h_t, mask_a, mask_t):
    """FusionWithEmotionDecoder.forward in train() mode with gradients enabled."""
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    names = tuple(n for n, _ in named)
    return _FusionFunction.apply(model, names, h_a, h_t, mask_a, mask_t, *[p for _, p in named])


class _MoseiFunction(torch.autograd.Function):
    """MoseiFusionWithEmotionDecoder: audio_proj / text_proj (models/mosei_fusion_with_emotion_decoder.py:64-66) in front
    of the backbone; their weight gradients are dX_backbone^T . features (tcgen05 wgrad on the zero-padded bf16 inputs)."""

    @staticmethod
    def forward(ctx, wrapper, names, h_a, h_t, mask_a, mask_t, *params):
        with torch.no_grad():
            xa, xt, a_in, t_in = wrapper.project_inputs(h_a, h_t, keep_inputs=True)
            logits, beta, z, tape = BW.forward_train(wrapper.backbone, xa, xt, mask_a, mask_t)
        ctx.wrapper, ctx.names, ctx.tape, ctx.inputs = wrapper, names, tape, (a_in, t_in)
        return logits, beta, z

    @staticmethod
    def backward(ctx, d_logits, d_beta, d_z):
        w = ctx.wrapper
        with torch.no_grad():
            grads, d_a, d_t = BW.backward_from(w.backbone, ctx.tape, d_logits, d_beta, d_z, need_dx=True)
            a_in, t_in = ctx.inputs
            full = {f"backbone.{k}": v for k, v in grads.items()}
            for name, d_x, x_in, lin in (("audio_proj", d_a, a_in, w.audio_proj), ("text_proj", d_t, t_in, w.text_proj)):
                dw, db = ops.linear_wgrad(d_x, x_in)                 # [d_model, K padded to a multiple of 8]
                full[f"{name}.weight"] = dw[:, : lin.weight.shape[1]].contiguous()
                full[f"{name}.bias"] = db
        ctx.tape = ctx.inputs = None
        params = dict(w.named_parameters())
        out = [None if n not in full else full[n].reshape(params[n].shape) for n in ctx.names]
        return (None, None, None, None, None, None) + tuple(out)


def mosei_forward_with_grad(wrapper, h_a, h_t, mask_a, mask_t):
    named = [(n, p) for n, p in wrapper.named_parameters() if p.requires_grad]
    names = tuple(n for n, _ in named)
    return _MoseiFunction.apply(wrapper, names, h_a, h_t, mask_a, mask_t, *[p for _, p in named])
