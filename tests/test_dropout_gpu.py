"""Dropout of the training step on a B200 (hriemo/dropout.py, csrc/dropout.cuh): the reference applies nn.Dropout at
every sub-layer output and nn.MultiheadAttention(dropout=p) on the attention probabilities
(models/cross_modal_block_tacfn.py:24-38, 81-119; models/emotion_decoder.py:14-29, 42-59).

The masks are counter-based, so every kernel can be checked EXACTLY against torch with the very mask it used: the mask
kernel is pinned to the torch restatement of the hash, the attention kernels (tcgen05 forward and backward, the decoder's)
are compared with (softmax(S) o M / (1 - p)) V and its autograd, and the whole training step on the GPU is compared with
the float64 CPU schedule (tests/kernel_standins.py) run with the same stream keys."""
import math

import pytest
import torch

import kernel_standins

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(got, ref):
    ref = ref.double().to(got.device)
    return ((got.double() - ref).norm() / (ref.norm() + 1e-300)).item()


def test_mask_kernel_matches_the_torch_restatement_of_the_hash():
    from hriemo import dropout as D, ops

    key, p8 = D.Drop(0.2, 7).site(1005)[2], 51
    m = ops.dropout_mask(300, 770, key, p8, DEV)
    assert torch.equal(m.cpu(), D.keep_mask(300, 770, key, p8))
    assert abs(m.double().mean().item() - (1 - 51 / 256)) < 4e-3
    # attention streams: rows_per_stream = Tq, stream s = b * H + h under drop_key_bh
    B, H, Tq, Tk = 2, 3, 37, 53
    ma = ops.dropout_mask(B * H * Tq, Tk, key, p8, DEV, rows_per_stream=Tq).view(B * H, Tq, Tk).cpu()
    for s in range(B * H):
        assert torch.equal(ma[s], D.keep_mask(Tq, Tk, D.key_bh(key, s), p8))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_dropout_kernel_forward_and_backward_use_one_mask(dtype):
    from hriemo import dropout as D, ops

    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(1000, 768, device=DEV, generator=g).to(dtype)
    r = torch.randn(1000, 768, device=DEV, generator=g).to(dtype)
    drop = D.Drop(0.1, 12345).site(2007)
    p8, scale, key = drop
    keep = D.keep_mask(1000, 768, key, p8).to(DEV)
    y = ops.dropout(x, drop, resid=r)
    ref = x.float() * keep * scale + r.float()
    assert (y.float() - ref).abs().max().item() <= (2e-2 if dtype == torch.bfloat16 else 1e-6) * ref.abs().max().item()
    dy = ops.dropout(x, drop)      # the backward's call: same key, no residual
    assert torch.equal(dy == 0, ~keep | (x == 0))
    assert (dy.float() - x.float() * keep * scale).abs().max().item() <= (2e-2 if dtype == torch.bfloat16 else 1e-6) * 5.0
    # a strided view (column slice of a packed buffer) is a legal operand
    wide = torch.randn(64, 1536, device=DEV, generator=g).to(dtype)
    part = ops.dropout(wide[:, 768:], drop)
    assert torch.equal(part, ops.dropout(wide[:, 768:].contiguous(), drop))


def _attn_ref(q, k, v, pad, keep, scale_drop, B, H, Tq, Tk, dh):
    qh = q.view(B, Tq, H, dh).transpose(1, 2)
    kh = k.view(B, Tk, H, dh).transpose(1, 2)
    vh = v.view(B, Tk, H, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    if pad is not None:
        s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1) * keep * scale_drop
    return (p @ vh).transpose(1, 2).reshape(B * Tq, H * dh)


@pytest.mark.parametrize("B,H,Tq,Tk,dh,masked", [(3, 2, 301, 299, 96, True), (2, 4, 64, 300, 96, False), (2, 2, 300, 64, 96, True),
                                                  (2, 4, 128, 128, 64, True), (2, 2, 500, 500, 96, False)])
def test_attention_forward_and_backward_with_probability_dropout(B, H, Tq, Tk, dh, masked):
    """The tcgen05 kernels (general, paired-head and DEFER forms of the forward; both passes of the backward) against
    torch with the mask they used."""
    from hriemo import dropout as D, ops

    g = torch.Generator(device=DEV).manual_seed(B * 1000 + Tq + Tk)
    d = H * dh
    q = torch.randn(B * Tq, d, device=DEV, generator=g).bfloat16()
    k = torch.randn(B * Tk, d, device=DEV, generator=g).bfloat16()
    v = torch.randn(B * Tk, d, device=DEV, generator=g).bfloat16()
    d_out = torch.randn(B * Tq, d, device=DEV, generator=g).bfloat16()
    pad = None
    if masked:
        lens = torch.randint(Tk // 3, Tk + 1, (B, 1), device=DEV, generator=g)
        pad = torch.arange(Tk, device=DEV)[None, :] >= lens
    drop = D.Drop(0.15, 4242).site(1005)
    p8, sc, key = drop
    keep = ops.dropout_mask(B * H * Tq, Tk, key, p8, DEV, rows_per_stream=Tq).view(B, H, Tq, Tk).double()
    out, lse = ops.attention(q, k, v, pad, B, H, Tq, Tk, dh, want_lse=True, drop=drop)
    out0, lse0 = ops.attention(q, k, v, pad, B, H, Tq, Tk, dh, want_lse=True)
    assert torch.equal(lse, lse0)                      # the row statistics are those of the undropped probabilities
    assert not torch.equal(out, out0)
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qd, kd, vd, pad, keep, sc, B, H, Tq, Tk, dh)
    assert _rel(out, ref.detach()) <= 1e-2
    ref.backward(d_out.double())
    dq, dk, dv = ops.attention_backward(q, k, v, out, d_out, lse, pad, B, H, Tq, Tk, dh, drop=drop)
    for name, got, want in (("dq", dq, qd.grad), ("dk", dk, kd.grad), ("dv", dv, vd.grad)):
        assert _rel(got, want) <= 1.5e-2, name


@pytest.mark.parametrize("B,H,Nq,Tk,dh,masked", [(5, 2, 4, 64, 96, True), (3, 4, 6, 128, 64, False), (4, 8, 4, 4, 96, False)])
def test_decoder_attention_with_probability_dropout(B, H, Nq, Tk, dh, masked):
    from hriemo import dropout as D, ops

    g = torch.Generator(device=DEV).manual_seed(Nq * 100 + Tk)
    d = H * dh
    q = torch.randn(B * Nq, d, device=DEV, generator=g).bfloat16()
    kv = torch.randn(B * Tk, 2 * d, device=DEV, generator=g).bfloat16()
    d_out = torch.randn(B * Nq, d, device=DEV, generator=g).bfloat16()
    pad = None
    if masked:
        lens = torch.randint(1, Tk + 1, (B, 1), device=DEV, generator=g)
        pad = torch.arange(Tk, device=DEV)[None, :] >= lens
    drop = D.Drop(0.2, 99).site(100003)
    p8, sc, key = drop
    keep = ops.dropout_mask(B * H * Nq, Tk, key, p8, DEV, rows_per_stream=Nq).view(B, H, Nq, Tk).double()
    out, _ = ops.small_attention(q, kv[:, :d], kv[:, d:], pad, B, H, Nq, Tk, dh, drop=drop)
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, kv[:, :d], kv[:, d:]))
    ref = _attn_ref(qd, kd, vd, pad, keep, sc, B, H, Nq, Tk, dh)
    assert _rel(out, ref.detach()) <= 6e-3
    ref.backward(d_out.double())
    dq, dk, dv = ops.small_attention_backward(q, kv[:, :d], kv[:, d:], d_out, pad, B, H, Nq, Tk, dh, drop=drop)
    for name, got, want in (("dq", dq, qd.grad), ("dk", dk, kd.grad), ("dv", dv, vd.grad)):
        assert _rel(got, want) <= 1e-2, name
    with pytest.raises(Exception, match="attention maps"):
        ops.small_attention(q, kv[:, :d], kv[:, d:], pad, B, H, Nq, Tk, dh, want_probs=True, drop=drop)


def test_training_step_with_dropout_matches_the_float64_schedule_with_the_same_keys(monkeypatch):
    """Whole model, p = 0.1, ragged masks: loss, logits and all 119 gradients of backward.loss_and_gradients on the GPU
    against the float64 CPU schedule (every kernel a torch stand-in, masks from the torch restatement of the hash) run
    with the same stream keys -- every dropout site of the forward and of the backward has to agree."""
    import copy

    from hriemo import backward, dropout as D
    from hriemo.train import invalidate_prepared
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    B, T_a, T_t, d, Ne = 6, 70, 24, 256, 4
    torch.manual_seed(91)
    model = FusionWithEmotionDecoder(d_model=d, n_heads=4, num_emotions=Ne, beta_hidden=64, dropout=0.1).train()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    ref_model = copy.deepcopy(model).double()
    g = torch.Generator().manual_seed(92)
    h_a = torch.randn(B, T_a, d, generator=g).bfloat16().float()
    h_t = torch.randn(B, T_t, d, generator=g).bfloat16().float()
    ma = torch.arange(T_a)[None, :] >= torch.randint(T_a // 2, T_a + 1, (B, 1), generator=g)
    mt = torch.arange(T_t)[None, :] >= torch.randint(T_t // 2, T_t + 1, (B, 1), generator=g)
    labels = torch.eye(Ne)[torch.randint(0, Ne, (B,), generator=g)]
    monkeypatch.setattr(D, "make", lambda p: D.Drop(p, 777) if p > 0 else None)
    model = model.to(DEV)
    out = backward.loss_and_gradients(model, h_a.to(DEV), h_t.to(DEV), ma.to(DEV), mt.to(DEV), labels.to(DEV))
    torch.cuda.synchronize()
    # and dropout is really on: the eval()-mode pass of the same function is the p = 0 step
    model.eval()
    out_eval = backward.loss_and_gradients(model, h_a.to(DEV), h_t.to(DEV), ma.to(DEV), mt.to(DEV), labels.to(DEV))
    assert (out_eval["logits"] - out["logits"]).abs().max().item() > 1e-2
    got = {k: v.detach().double().cpu() for k, v in out["grads"].items()}
    got_loss, got_logits = out["loss"].item(), out["logits"].double().cpu()

    kernel_standins.install(monkeypatch, exact=True)
    monkeypatch.setattr(backward.E, "to_seq", lambda x, what, ld=None: backward.E.Seq(x.reshape(-1, x.shape[-1]), x.shape[0], x.shape[1]))
    invalidate_prepared(ref_model)
    ref = backward.loss_and_gradients(ref_model, h_a.double(), h_t.double(), ma, mt, labels.double())
    assert abs(got_loss - ref["loss"].item()) <= 1e-2
    assert (got_logits - ref["logits"]).abs().max().item() <= 6e-2
    assert set(got) == set(ref["grads"])
    errs = {k: _rel(got[k], ref["grads"][k]) for k in got}
    bad = {k: v for k, v in errs.items() if not v <= (1.5e-1 if (".linear1." in k or ".ffn_a.0." in k or ".ffn_t.0." in k) else 6e-2)}
    assert not bad, f"relative errors out of bounds: {bad}"


def test_trainer_and_autograd_boundary_apply_dropout_in_train_mode():
    from hriemo.train import Trainer
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(5)
    model = FusionWithEmotionDecoder(d_model=256, n_heads=4, num_emotions=4, beta_hidden=64, dropout=0.2).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(6)
    B, T_a, T_t = 8, 50, 20
    h_a, h_t = torch.randn(B, T_a, 256, device=DEV, generator=g), torch.randn(B, T_t, 256, device=DEV, generator=g)
    y = (torch.rand(B, 4, device=DEV, generator=g) < 0.5).float()
    # the reference's loop through the autograd boundary: two forwards in train() mode draw different masks; the same
    # torch seed reproduces them; eval() is deterministic and different
    model.train()
    torch.manual_seed(100)
    lo1 = model(h_a, h_t)[0].detach().clone()
    lo2 = model(h_a, h_t)[0].detach().clone()
    torch.manual_seed(100)
    lo1b = model(h_a, h_t)[0].detach().clone()
    assert torch.equal(lo1, lo1b) and not torch.equal(lo1, lo2)
    logits, beta, _ = model(h_a, h_t)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    model.eval()
    with torch.no_grad():
        e1, e2 = model(h_a, h_t)[0], model(h_a, h_t)[0]
    assert torch.equal(e1, e2) and (e1 - lo1).abs().max().item() > 1e-3
    # Trainer: steps in train() mode run (eagerly) with dropout and stay finite
    model.train()
    with pytest.warns(UserWarning, match="CUDA-graph replay"):
        tr = Trainer(model, graph=True, distributed=False)
    for _ in range(4):
        info = tr.step(h_a, h_t, None, None, y)
    assert torch.isfinite(info["loss"]).all() and torch.isfinite(info["grad_norm"]).all() and tr._graph is None
