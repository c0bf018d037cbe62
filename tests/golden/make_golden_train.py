#!/usr/bin/env python3
"""Generates tests/golden/train_step_*.pt: two consecutive optimizer steps of the UNMODIFIED reference
(imported read-only from /root/reference), following scripts/fusion/train_fusion_seq_level_decoder.py:300-339
with the setup of :395-416 (AdamW lr 1e-4, weight decay 1e-2, BCEWithLogitsLoss, clip 5.0) and dropout = 0.
Run in the authoring container:   python tests/golden/make_golden_train.py

The 54.6 M-parameter gradients are not stored: a fixture holds the loss, the logits and beta, the total
gradient norm, the L2 norm of every parameter's gradient and of every parameter's update, and three
gradients in full; weights are rebuilt from the seed by the tests (checksum-verified)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import checksum, inputs, ref, save  # noqa: E402

FULL = ["emotion_decoder.out_proj.weight", "emotion_decoder.emotion_queries", "beta_gate.mlp.2.bias"]


def train_case(name, ctor, model_seed, in_seed, B, T_a, T_t):
    torch.manual_seed(model_seed)
    model = ref("fusion_with_emotion_decoder").FusionWithEmotionDecoder(dropout=0.0, **ctor)
    model.train()
    weights = checksum(model.state_dict())
    d = ctor.get("d_model", 768)
    n_e = ctor.get("num_emotions", 4)
    h_a, h_t, m_a, m_t = inputs(in_seed, B, T_a, T_t, d, d, True)
    g = torch.Generator().manual_seed(in_seed + 1)
    labels = torch.eye(n_e)[torch.randint(0, n_e, (B,), generator=g)]          # one-hot multi-label rows (:164-171)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-2)
    crit = torch.nn.BCEWithLogitsLoss()
    names = [k for k, _ in model.named_parameters()]
    full = [k for k in FULL if k in names]
    steps = []
    for _ in range(2):
        before = {k: p.detach().clone() for k, p in model.named_parameters()}
        logits, beta, _ = model(h_a, h_t, m_a, m_t)
        loss = crit(logits, labels)
        loss = loss - 0.01 * (beta * (1 - beta)).mean()
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        total = float(torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0))
        opt.step()
        opt.zero_grad()
        steps.append(dict(loss=float(loss), logits=logits.detach().clone(), beta=beta.detach().clone(), grad_norm=total,
                          grad_norms={k: float(grads[k].double().norm()) for k in names},
                          update_norms={k: float((p.detach().double() - before[k].double()).norm())
                                        for k, p in model.named_parameters()},
                          grads_full={k: grads[k] for k in full},
                          params_full={k: dict(model.named_parameters())[k].detach().clone() for k in full}))
    save(name, dict(kind="train", ctor=ctor, model_seed=model_seed, in_seed=in_seed, B=B, T_a=T_a, T_t=T_t,
                    weights=weights, names=names, steps=steps, lr=1e-4, weight_decay=1e-2, max_norm=5.0))


if __name__ == "__main__":
    torch.backends.mha.set_fastpath_enabled(False)
    train_case("train_step_default", {}, 1234, 31, 4, 60, 20)
    train_case("train_step_small", dict(d_model=192, n_heads=2, beta_hidden=64, num_emotions=5), 7, 32, 6, 90, 30)
