"""CPU oracle for the HRI-EMO training step (BASELINE config 5).

TEST INFRASTRUCTURE ONLY, like hriemo_oracle.py: nothing in the product package may import it.

What it restates
----------------
One iteration of the reference's train_one_epoch
(scripts/fusion/train_fusion_seq_level_decoder.py:300-339, setup :405-416):

    logits, beta, _ = model(h_a, h_t, m_a, m_t)                  :311
    loss = BCEWithLogitsLoss()(logits, labels)                   :319, :416   (multi-label)
    loss = loss - 0.01 * (beta * (1 - beta)).mean()              :326-327
    loss.backward()                                              :332
    clip_grad_norm_(model.parameters(), max_norm=5.0)            :333
    AdamW(lr, weight_decay).step(); zero_grad()                  :334-335, :405-409

The forward is the functional oracle of hriemo_oracle.py (dropout = 0, the configuration in which the
reference is deterministic and parity is defined).  The loss, the global-norm clip
and the AdamW update are restated here from their published definitions (PyTorch's, which the reference
calls and does not vendor); the backward pass is reverse-mode differentiation of the restated forward,
carried out by torch.autograd over the oracle's plain tensor arithmetic -- the same engine the reference's
`loss.backward()` uses, applied to independently written forward code, so agreement of the gradients checks
the forward restatement's derivative, not autograd against itself.

Parity pin: tests/golden/train_step_*.pt, produced by tests/golden/make_golden_train.py from the imported
reference (two consecutive optimizer steps: loss, gradient norms per parameter, clip coefficient, parameter
updates, and three gradients in full); checked by tests/test_oracle_train.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

import hriemo_oracle as O

Tensor = torch.Tensor
State = Dict[str, Tensor]


def bce_with_logits(logits: Tensor, targets: Tensor) -> Tensor:
    """nn.BCEWithLogitsLoss(reduction="mean"): mean over every element of
    max(x, 0) - x*y + log(1 + exp(-|x|))   (the numerically stable form of -[y log s(x) + (1-y) log(1-s(x))])."""
    x, y = logits, targets
    return (x.clamp(min=0) - x * y + torch.log1p(torch.exp(-x.abs()))).mean()


def beta_regulariser(beta: Tensor) -> Tensor:
    """(beta * (1 - beta)).mean(), subtracted from the loss with weight 0.01 (:326-327)."""
    return (beta * (1 - beta)).mean()


def train_loss(sd: State, h_a: Tensor, h_t: Tensor, m_a, m_t, labels: Tensor, n_heads: int = 8,
               beta_weight: float = 0.01) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (loss, logits, beta) of one forward in training mode with dropout = 0."""
    logits, beta, _ = O.fusion_with_emotion_decoder(sd, h_a, h_t, m_a, m_t, n_heads=n_heads)
    loss = bce_with_logits(logits, labels) - beta_weight * beta_regulariser(beta)
    return loss, logits, beta


def clip_coefficient(grads: Dict[str, Tensor], max_norm: float = 5.0) -> Tuple[float, float]:
    """torch.nn.utils.clip_grad_norm_ (L2): total = sqrt(sum_p ||g_p||^2); every gradient is multiplied by
    min(1, max_norm / (total + 1e-6)).  Returns (total, coefficient)."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values()))
    return total, min(1.0, max_norm / (total + 1e-6))


def adamw_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float = 1e-4, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2) -> Tuple[Tensor, Tensor, Tensor]:
    """torch.optim.AdamW, step `step` (1-based), decoupled weight decay:
        p <- p * (1 - lr * wd);  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2;
        p <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)."""
    b1, b2 = betas
    p = p * (1.0 - lr * weight_decay)
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    denom = v.sqrt() / math.sqrt(1.0 - b2 ** step) + eps
    p = p - (lr / (1.0 - b1 ** step)) * m / denom
    return p, m, v


def train_step(sd: State, opt: Optional[dict], h_a: Tensor, h_t: Tensor, m_a, m_t, labels: Tensor, n_heads: int = 8,
               lr: float = 1e-4, weight_decay: float = 1e-2, max_norm: float = 5.0, beta_weight: float = 0.01):
    """One optimizer step.  sd: parameter name -> tensor (not modified); opt: {"step", "m", "v"} or None for a
    fresh optimizer.  Returns (new_sd, new_opt, info) with info = loss, logits, beta, grads (unclipped),
    grad_norm (total, before clipping) and clip (the coefficient applied)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    loss, logits, beta = train_loss(params, h_a, h_t, m_a, m_t, labels, n_heads, beta_weight)
    names = list(params)
    gs = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(params[k])) for k, g in zip(names, gs)}
    total, coef = clip_coefficient(grads, max_norm)
    step = 1 if opt is None else opt["step"] + 1
    new_sd, new_m, new_v = {}, {}, {}
    for k in names:
        m = torch.zeros_like(sd[k]) if opt is None else opt["m"][k]
        v = torch.zeros_like(sd[k]) if opt is None else opt["v"][k]
        new_sd[k], new_m[k], new_v[k] = adamw_update(sd[k].detach(), grads[k] * coef, m, v, step, lr=lr,
                                                     weight_decay=weight_decay)
    info = dict(loss=float(loss.detach()), logits=logits.detach(), beta=beta.detach(), grads=grads, grad_norm=total, clip=coef)
    return new_sd, dict(step=step, m=new_m, v=new_v), info
