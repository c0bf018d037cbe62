"""The schedule of hriemo/backward.py on the CPU: every kernel wrapper is replaced by a torch stand-in
(tests/kernel_standins.py, test infrastructure) and the composed gate -> decoder -> loss backward is compared with
autograd over the oracle's float64 forward.  This pins the host logic — saved activations, operand order,
residual joins, parameter names — without a GPU; the kernels are checked on the B200 (tests/test_backward_gpu.py).
Two modes: exact (all stand-ins in float64 without any rounding: the schedule must reproduce autograd to 1e-9,
which is the logic check) and bf16 (stand-ins round where the kernels do: shows the noise level of bf16 activations
and activation gradients on tiny batches, where a ReLU whose pre-activation changes sign under bf16 rounding costs
a whole element of the gradient; bound 0.2 on the whole-tensor relative error)."""
import pytest
import torch

import kernel_standins
from test_backward_gpu import _oracle_backward, _rel


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("B,T_a,T_t,d,H,Ne,masked", [(6, 20, 12, 256, 4, 4, True), (5, 9, 9, 128, 2, 6, False)])
def test_decode_loss_and_backward_schedule(monkeypatch, B, T_a, T_t, d, H, Ne, masked, exact):
    kernel_standins.install(monkeypatch, exact=exact)
    tol = 1e-9 if exact else 0.2
    from hriemo import backward, engine as E
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(7)
    model = FusionWithEmotionDecoder(d_model=d, num_emotions=Ne, n_heads=H, num_layers_fusion=1, num_layers_decoder=2,
                                     beta_hidden=64, dropout=0.0)
    if exact:
        model = model.double()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    g = torch.Generator().manual_seed(8)
    a = torch.randn(B, T_a, d, generator=g).bfloat16()
    t = torch.randn(B, T_t, d, generator=g).bfloat16()
    if exact:
        a, t = a.double(), t.double()
    ma = mt = None
    if masked:
        la = torch.randint(1, T_a + 1, (B,), generator=g)
        lt = torch.randint(1, T_t + 1, (B,), generator=g)
        ma = torch.arange(T_a)[None, :] >= la[:, None]
        mt = torch.arange(T_t)[None, :] >= lt[:, None]
    labels = (torch.rand(B, Ne, generator=g) < 0.4).to(torch.float64 if exact else torch.float32)

    out = backward.decode_loss_and_backward(model, E.Seq(a.view(B * T_a, d), B, T_a), E.Seq(t.view(B * T_t, d), B, T_t),
                                            ma, mt, labels)
    loss, logits, beta, grads, d_a, d_t = _oracle_backward(model, a, t, ma, mt, labels, H)
    assert abs(out["loss"].item() - loss.item()) <= (1e-12 if exact else 5e-3)
    assert (out["logits"].double() - logits).abs().max().item() <= (1e-10 if exact else 3e-2)
    assert set(out["grads"]) == set(grads), sorted(set(out["grads"]) ^ set(grads))
    errs = {"d_a": _rel(out["d_a"], d_a.view(B * T_a, d)), "d_t": _rel(out["d_t"], d_t.view(B * T_t, d))}
    for k, ref in grads.items():
        assert tuple(out["grads"][k].shape) == tuple(ref.shape), k
        errs[k] = _rel(out["grads"][k], ref)
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"relative errors above {tol}: {bad}"


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("B,T_a,T_t,d,H,Ne,masked", [(4, 14, 9, 256, 4, 4, True), (3, 8, 8, 128, 2, 6, False)])
def test_loss_and_gradients_schedule_whole_model(monkeypatch, B, T_a, T_t, d, H, Ne, masked, exact):
    """Every one of the model's parameters (two encoder layers, gate, two decoder layers) against autograd over the
    oracle's forward: 1e-9 in float64, the bf16 noise bound otherwise."""
    import hriemo_oracle_train as OT

    kernel_standins.install(monkeypatch, exact=exact)
    from hriemo import backward
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(11)
    model = FusionWithEmotionDecoder(d_model=d, num_emotions=Ne, n_heads=H, num_layers_fusion=2, num_layers_decoder=2,
                                     beta_hidden=64, dropout=0.0)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
            if n.endswith("in_proj_bias") or n.endswith("out_proj.bias"):
                p.add_(0.05 * torch.randn_like(p))
    dt = torch.float64 if exact else torch.float32
    if exact:
        model = model.double()
    g = torch.Generator().manual_seed(12)
    h_a = torch.randn(B, T_a, d, generator=g).bfloat16().to(dt)
    h_t = torch.randn(B, T_t, d, generator=g).bfloat16().to(dt)
    ma = mt = None
    if masked:
        ma = torch.arange(T_a)[None, :] >= torch.randint(1, T_a + 1, (B,), generator=g)[:, None]
        mt = torch.arange(T_t)[None, :] >= torch.randint(1, T_t + 1, (B,), generator=g)[:, None]
    labels = (torch.rand(B, Ne, generator=g) < 0.4).to(dt)
    if exact:   # to_seq casts fp32 features to bf16; in float64 mode hand the streams over as they are
        monkeypatch.setattr(backward.E, "to_seq", lambda x, what, ld=None: backward.E.Seq(x.reshape(-1, x.shape[-1]), x.shape[0], x.shape[1]))
    out = backward.loss_and_gradients(model, h_a, h_t, ma, mt, labels)

    sd = {k: v.detach().double().requires_grad_(True) for k, v in model.state_dict().items()}
    loss, _, _ = OT.train_loss(sd, h_a.double(), h_t.double(), ma, mt, labels.double(), n_heads=H)
    loss.backward()
    assert abs(out["loss"].item() - loss.item()) <= (1e-12 if exact else 5e-3)
    assert set(out["grads"]) == set(sd), sorted(set(out["grads"]) ^ set(sd))
    assert set(out["grads"]) == {n for n, _ in model.named_parameters()}
    tol = 1e-9 if exact else 0.25
    errs = {}
    for k, p in sd.items():
        assert tuple(out["grads"][k].shape) == tuple(p.shape), k
        errs[k] = _rel(out["grads"][k], p.grad)
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"relative errors above {tol}: {bad}"


def test_trainer_two_steps_match_the_training_oracle(monkeypatch):
    """hriemo.train.Trainer (flat arenas, clip, AdamW, prepared-operand invalidation) over the float64 stand-ins
    against two consecutive steps of oracle/hriemo_oracle_train.py (itself pinned to the reference's steps)."""
    import hriemo_oracle_train as OT

    kernel_standins.install(monkeypatch, exact=True)
    from hriemo import backward
    from hriemo.train import Trainer
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    B, T_a, T_t, d, H, Ne = 3, 10, 6, 128, 2, 4
    torch.manual_seed(21)
    model = FusionWithEmotionDecoder(d_model=d, num_emotions=Ne, n_heads=H, num_layers_fusion=1, num_layers_decoder=1,
                                     beta_hidden=32, dropout=0.0).double()
    g = torch.Generator().manual_seed(22)
    h_a = torch.randn(B, T_a, d, generator=g, dtype=torch.float64)
    h_t = torch.randn(B, T_t, d, generator=g, dtype=torch.float64)
    labels = torch.eye(Ne, dtype=torch.float64)[torch.randint(0, Ne, (B,), generator=g)]
    monkeypatch.setattr(backward.E, "to_seq", lambda x, what, ld=None: backward.E.Seq(x.reshape(-1, x.shape[-1]), x.shape[0], x.shape[1]))
    sd = {k: p.detach().clone() for k, p in model.named_parameters()}
    trainer = Trainer(model, lr=1e-3, max_norm=0.5, distributed=False)
    assert all(p.data_ptr() >= trainer.params.data_ptr() for p in model.parameters())
    opt = None
    for _ in range(2):
        info = trainer.step(h_a, h_t, None, None, labels)
        sd, opt, want = OT.train_step(sd, opt, h_a, h_t, None, None, labels, n_heads=H, lr=1e-3, max_norm=0.5)
        assert abs(info["loss"].item() - want["loss"]) <= 1e-12
        assert abs(info["grad_norm"].item() - want["grad_norm"]) <= 1e-9 * want["grad_norm"]
        assert abs(info["clip"].item() - want["clip"]) <= 1e-9
        for k, p in model.named_parameters():
            assert (p.detach() - sd[k]).abs().max().item() <= 1e-9, k
    assert want["clip"] < 1.0   # the clip was active


def test_trainer_keeps_the_state_dict_contract(monkeypatch):
    """Moving the parameters into the flat arena leaves state_dict() keys / shapes / values as they were, and a strict
    load_state_dict() afterwards writes INTO the arena (the modules' tensors stay views of it)."""
    kernel_standins.install(monkeypatch, exact=True)
    from hriemo.train import Trainer
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(41)
    model = FusionWithEmotionDecoder(d_model=128, num_emotions=4, n_heads=2, num_layers_fusion=1, num_layers_decoder=1,
                                     beta_hidden=32, dropout=0.0).double()
    before = {k: v.clone() for k, v in model.state_dict().items()}
    trainer = Trainer(model, distributed=False)
    after = model.state_dict()
    assert list(after) == list(before)
    assert all(torch.equal(after[k], before[k]) for k in before)
    assert trainer.numel >= sum(v.numel() for v in before.values())
    lo, hi = trainer.params.data_ptr(), trainer.params.data_ptr() + trainer.params.numel() * trainer.params.element_size()
    assert all(lo <= p.data_ptr() < hi and p.data_ptr() % 16 == 0 for p in model.parameters())
    other = {k: torch.randn_like(v) for k, v in before.items()}
    model.load_state_dict(other, strict=True)
    assert all(lo <= p.data_ptr() < hi for p in model.parameters())
    for name, (o, n) in trainer.slots.items():
        assert torch.equal(trainer.params[o:o + n], other[name].reshape(-1)), name
    # dropout > 0 is applied by the training step in train() mode; only its combination with CUDA-graph replay (the mask
    # keys are host scalars baked into a captured graph) is announced and falls back to eager steps
    with pytest.warns(UserWarning, match="CUDA-graph replay"):
        Trainer(FusionWithEmotionDecoder(d_model=128, num_emotions=4, n_heads=2, num_layers_fusion=1,
                                         num_layers_decoder=1, beta_hidden=32, dropout=0.1).double(), distributed=False, graph=True)


@pytest.mark.parametrize("L_f,L_d,Ne,mask_mode", [(3, 1, 8, "audio"), (1, 3, 1, "text"), (0, 2, 4, "both")])
def test_loss_and_gradients_schedule_other_depths_and_masks(monkeypatch, L_f, L_d, Ne, mask_mode):
    """Exact (float64) schedule check on the shapes the other tests do not visit: three encoder layers / none at all,
    one and three decoder layers, one and eight emotion queries, a PAD mask on one modality only, T_a == T_t."""
    import hriemo_oracle_train as OT

    kernel_standins.install(monkeypatch, exact=True)
    from hriemo import backward
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    B, T, d, H = 3, 7, 128, 4
    torch.manual_seed(51)
    model = FusionWithEmotionDecoder(d_model=d, num_emotions=Ne, n_heads=H, num_layers_fusion=L_f, num_layers_decoder=L_d,
                                     beta_hidden=32, dropout=0.0).double()
    g = torch.Generator().manual_seed(52)
    h_a = torch.randn(B, T + (0 if mask_mode == "both" else 4), d, generator=g, dtype=torch.float64)
    h_t = torch.randn(B, T, d, generator=g, dtype=torch.float64)
    ma = torch.arange(h_a.shape[1])[None, :] >= torch.randint(1, h_a.shape[1] + 1, (B,), generator=g)[:, None]
    mt = torch.arange(T)[None, :] >= torch.randint(1, T + 1, (B,), generator=g)[:, None]
    if mask_mode == "audio":
        mt = None
    if mask_mode == "text":
        ma = None
    labels = (torch.rand(B, Ne, generator=g) < 0.5).double()
    monkeypatch.setattr(backward.E, "to_seq", lambda x, what, ld=None: backward.E.Seq(x.reshape(-1, x.shape[-1]), x.shape[0], x.shape[1]))
    out = backward.loss_and_gradients(model, h_a, h_t, ma, mt, labels)
    sd = {k: v.detach().double().requires_grad_(True) for k, v in model.state_dict().items()}
    loss, _, _ = OT.train_loss(sd, h_a, h_t, ma, mt, labels, n_heads=H)
    loss.backward()
    assert abs(out["loss"].item() - loss.item()) <= 1e-12
    assert set(out["grads"]) == set(sd)
    bad = {k: _rel(out["grads"][k], p.grad) for k, p in sd.items() if not _rel(out["grads"][k], p.grad) <= 1e-9}
    assert not bad, bad


def test_trainer_gradient_accumulation_and_lr_schedule(monkeypatch):
    """step_accumulated over two equal micro-batches equals one oracle step on their union (the loss is a mean over
    utterances), and `trainer.lr` set between steps is the learning rate AdamW uses (float64 stand-ins, 1e-9)."""
    import hriemo_oracle_train as OT

    kernel_standins.install(monkeypatch, exact=True)
    from hriemo import backward
    from hriemo.train import Trainer
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    B, T_a, T_t, d, H, Ne = 4, 9, 5, 128, 2, 4
    torch.manual_seed(61)
    model = FusionWithEmotionDecoder(d_model=d, num_emotions=Ne, n_heads=H, num_layers_fusion=1, num_layers_decoder=1,
                                     beta_hidden=32, dropout=0.0).double()
    g = torch.Generator().manual_seed(62)
    h_a = torch.randn(B, T_a, d, generator=g, dtype=torch.float64)
    h_t = torch.randn(B, T_t, d, generator=g, dtype=torch.float64)
    labels = (torch.rand(B, Ne, generator=g) < 0.5).double()
    monkeypatch.setattr(backward.E, "to_seq", lambda x, what, ld=None: backward.E.Seq(x.reshape(-1, x.shape[-1]), x.shape[0], x.shape[1]))
    sd = {k: p.detach().clone() for k, p in model.named_parameters()}
    trainer = Trainer(model, lr=1e-3, max_norm=0.5, distributed=False)
    opt = None
    for lr in (1e-3, 4e-4):
        trainer.lr = lr
        info = trainer.step_accumulated([(h_a[:2], h_t[:2], None, None, labels[:2]), (h_a[2:], h_t[2:], None, None, labels[2:])])
        sd, opt, want = OT.train_step(sd, opt, h_a, h_t, None, None, labels, n_heads=H, lr=lr, max_norm=0.5)
        assert abs(info["loss"].item() - want["loss"]) <= 1e-12
        assert abs(info["grad_norm"].item() - want["grad_norm"]) <= 1e-9 * want["grad_norm"]
        for k, p in model.named_parameters():
            assert (p.detach() - sd[k]).abs().max().item() <= 1e-9, k


def test_dropout_schedule_is_the_gradient_of_its_forward(monkeypatch):
    """Dropout > 0 (hriemo/dropout.py): with the mask keys of a step held fixed, the training forward is a smooth
    function of the parameters and the hand-scheduled backward must be its gradient -- every one of the 16 dropout
    sites of a layer (sub-layer outputs, the decoder FFN's inner dropout, the probabilities of all six MHAs) has to be
    re-applied in the backward at the right place with the same stream key.  Exact float64 stand-ins (whose masks are
    the arithmetic of csrc/dropout.cuh restated in torch); central differences on entries of every parameter family."""
    kernel_standins.install(monkeypatch, exact=True)
    from hriemo import backward, dropout as D
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(11)
    model = FusionWithEmotionDecoder(d_model=128, num_emotions=3, n_heads=2, num_layers_fusion=2, num_layers_decoder=2,
                                     beta_hidden=32, dropout=0.25).double().train()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    g = torch.Generator().manual_seed(12)
    B, T_a, T_t, d = 3, 10, 6, 128
    a, t = torch.randn(B, T_a, d, generator=g).double(), torch.randn(B, T_t, d, generator=g).double()
    ma = torch.arange(T_a)[None, :] >= torch.tensor([10, 7, 4])[:, None]
    mt = torch.arange(T_t)[None, :] >= torch.tensor([6, 3, 5])[:, None]
    labels = (torch.rand(B, 3, generator=g) < 0.5).double()
    monkeypatch.setattr(D, "make", lambda p: D.Drop(p, 20240607) if p > 0 else None)   # the same masks for every call
    # to_seq casts fp32 features to bf16; in float64 mode hand the streams over as they are
    monkeypatch.setattr(backward.E, "to_seq", lambda x, what, ld=None: backward.E.Seq(x.reshape(-1, x.shape[-1]), x.shape[0], x.shape[1]))

    from hriemo.train import invalidate_prepared

    def loss_and_grads():
        invalidate_prepared(model)   # the entries below are perturbed through .data: the operand caches would not notice
        out = backward.loss_and_gradients(model, a, t, ma, mt, labels)
        return out["loss"].item(), out["grads"], out["logits"].clone()

    loss0, grads, logits_train = loss_and_grads()
    # dropout is really on: the eval-mode step differs, and is the p = 0 schedule
    model.eval()
    loss_eval, _, logits_eval = loss_and_grads()
    model.train()
    assert abs(loss_eval - loss0) > 1e-4 and (logits_eval - logits_train).abs().max().item() > 1e-3
    params = dict(model.named_parameters())
    gen = torch.Generator().manual_seed(13)
    checked = 0
    for name in ["cross_modal.layers.0.self_attn_a.in_proj_weight", "cross_modal.layers.0.self_attn_t.out_proj.weight",
                 "cross_modal.layers.0.attn_a2t.in_proj_weight", "cross_modal.layers.1.attn_t2a.out_proj.bias",
                 "cross_modal.layers.0.ffn_a.0.weight", "cross_modal.layers.1.ffn_t.2.weight", "cross_modal.layers.0.norm_a1.weight",
                 "cross_modal.layers.1.self_norm_t.bias", "beta_gate.mlp.0.weight", "emotion_decoder.emotion_queries",
                 "emotion_decoder.layers.0.self_attn.in_proj_weight", "emotion_decoder.layers.0.cross_attn.in_proj_weight",
                 "emotion_decoder.layers.1.cross_attn.out_proj.weight", "emotion_decoder.layers.0.linear1.weight",
                 "emotion_decoder.layers.1.linear2.weight", "emotion_decoder.layers.1.norm3.weight", "emotion_decoder.out_proj.weight"]:
        p = params[name]
        flat = p.data.view(-1)
        for idx in torch.randint(0, flat.numel(), (3,), generator=gen).tolist():
            old = flat[idx].item()
            eps = 1e-6
            flat[idx] = old + eps
            lp = loss_and_grads()[0]
            flat[idx] = old - eps
            lm = loss_and_grads()[0]
            flat[idx] = old
            num = (lp - lm) / (2 * eps)
            ana = grads[name].reshape(-1)[idx].item()
            assert abs(num - ana) <= 1e-6 + 1e-4 * abs(num), (name, idx, num, ana)
            checked += 1
    assert checked == 51


def test_dropout_mask_arithmetic_and_keys():
    """hriemo/dropout.py: keep rate of the quantised probability, determinism, distinct streams, p8 = 0 keeps everything."""
    from hriemo import dropout as D

    drop = D.Drop(0.1, 99)
    assert drop.p8 == 26 and abs(drop.scale - 1.0 / (1.0 - 26 / 256)) < 1e-12 and drop.on
    p8, scale, key = drop.site(1005)
    m = D.keep_mask(512, 768, key, p8)
    assert abs(m.double().mean().item() - (1 - 26 / 256)) < 3e-3
    assert torch.equal(m, D.keep_mask(512, 768, key, p8))
    other = D.keep_mask(512, 768, drop.site(1006)[2], p8)
    assert abs((m ^ other).double().mean().item() - 2 * (26 / 256) * (1 - 26 / 256)) < 5e-3     # independent streams
    assert D.keep_mask(8, 8, key, 0).all() and D.Drop(0.0, 1).site(1) is None and D.make(0.0) is None
    # rows and columns are decorrelated: neighbouring rows share no more than chance
    same = (m[1:] == m[:-1]).double().mean().item()
    assert abs(same - (1 - 2 * (26 / 256) * (1 - 26 / 256))) < 5e-3
    assert D.key_bh(key, 0) != D.key_bh(key, 1) and D.mix(0x12345678) == D.mix(0x12345678 + (1 << 32))
