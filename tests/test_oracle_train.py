"""Pins the training-step oracle (oracle/hriemo_oracle_train.py) against two consecutive optimizer steps of the
reference itself (tests/golden/train_step_*.pt from tests/golden/make_golden_train.py): loss, logits, beta, the
gradient norm of EVERY parameter, the clip coefficient, the update norm of every parameter after AdamW, and three
gradients / updated parameters in full.  The oracle runs in float64; the fixtures hold the reference's float32
results, so the bounds are float32 round-off of a backward pass through ~100 ops."""
import pytest
import torch

import golden_util as G
import hriemo_oracle as O
import hriemo_oracle_train as OT


def _build(fx):
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(fx["model_seed"])
    m = FusionWithEmotionDecoder(dropout=0.0, **fx["ctor"])
    G.assert_same_weights(m, fx["weights"])
    d, n_e = fx["ctor"].get("d_model", 768), fx["ctor"].get("num_emotions", 4)
    h_a, h_t, m_a, m_t = G.make_inputs(fx["in_seed"], fx["B"], fx["T_a"], fx["T_t"], d, d, True)
    g = torch.Generator().manual_seed(fx["in_seed"] + 1)
    labels = torch.eye(n_e)[torch.randint(0, n_e, (fx["B"],), generator=g)]
    return m, (h_a.double(), h_t.double(), m_a, m_t, labels.double())


@pytest.mark.parametrize("name", ["train_step_small", "train_step_default"])
def test_training_oracle_matches_two_reference_steps(name):
    fx = G.load(name)
    model, (h_a, h_t, m_a, m_t, labels) = _build(fx)
    assert [k for k, _ in model.named_parameters()] == fx["names"]
    sd = {k: p.detach().double() for k, p in model.named_parameters()}
    H = fx["ctor"].get("n_heads", 8)
    opt = None
    for step in fx["steps"]:
        new_sd, opt, info = OT.train_step(sd, opt, h_a, h_t, m_a, m_t, labels, n_heads=H, lr=fx["lr"],
                                          weight_decay=fx["weight_decay"], max_norm=fx["max_norm"])
        assert abs(info["loss"] - step["loss"]) <= 2e-6
        assert (info["logits"] - step["logits"]).abs().max().item() <= 2e-5
        assert (info["beta"] - step["beta"]).abs().max().item() <= 2e-5
        assert abs(info["grad_norm"] - step["grad_norm"]) <= 2e-4 * max(1.0, step["grad_norm"])
        assert info["clip"] == pytest.approx(min(1.0, fx["max_norm"] / (step["grad_norm"] + 1e-6)), rel=1e-3)
        worst = 0.0
        for k in fx["names"]:
            got, want = float(info["grads"][k].norm()), step["grad_norms"][k]
            worst = max(worst, abs(got - want) / max(want, 1e-3 * step["grad_norm"]))
        assert worst <= 2e-3, f"per-parameter gradient norm off by {worst:.2e} (relative)"
        for k, want in step["grads_full"].items():
            scale = max(float(want.abs().max()), 1e-6)
            assert (info["grads"][k] - want).abs().max().item() <= 2e-3 * scale, k
        for k in fx["names"]:
            got, want = float((new_sd[k] - sd[k]).norm()), step["update_norms"][k]
            # one AdamW step moves every element by ~lr (sign-like at step 1); float32 parameters quantise the
            # reference's update at ~6e-8 per element
            assert abs(got - want) <= 2e-2 * max(want, 1e-7) + 1e-6, (k, got, want)
        for k, want in step["params_full"].items():
            assert (new_sd[k] - want).abs().max().item() <= 5e-6, k
        sd = new_sd
    assert opt["step"] == 2
    assert info["clip"] < 1.0 or fx["steps"][-1]["grad_norm"] <= fx["max_norm"]


def test_loss_pieces_match_their_definitions():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(7, 5, generator=g, dtype=torch.float64) * 4
    y = (torch.rand(7, 5, generator=g) > 0.5).double()
    want = -(y * torch.log(torch.sigmoid(x)) + (1 - y) * torch.log(1 - torch.sigmoid(x))).mean()
    assert abs(float(OT.bce_with_logits(x, y)) - float(want)) <= 1e-12
    big = torch.tensor([[80.0, -80.0]], dtype=torch.float64)
    assert torch.isfinite(OT.bce_with_logits(big, torch.tensor([[0.0, 1.0]], dtype=torch.float64)))
    b = torch.tensor([[0.5], [0.0], [1.0]], dtype=torch.float64)
    assert float(OT.beta_regulariser(b)) == pytest.approx(0.25 / 3)
    total, coef = OT.clip_coefficient({"a": torch.full((4,), 3.0), "b": torch.full((9,), 4.0)}, max_norm=5.0)
    assert total == pytest.approx((36 + 144) ** 0.5) and coef == pytest.approx(5.0 / (total + 1e-6))
    p, m, v = OT.adamw_update(torch.ones(3, dtype=torch.float64), torch.full((3,), 0.5, dtype=torch.float64),
                              torch.zeros(3, dtype=torch.float64), torch.zeros(3, dtype=torch.float64), 1)
    # step 1: m_hat = g, v_hat = g^2 -> p = 1*(1 - lr*wd) - lr * g / (|g| + eps)
    assert float(p[0]) == pytest.approx(1.0 * (1 - 1e-4 * 1e-2) - 1e-4 * 0.5 / (0.5 + 1e-8), abs=1e-12)
