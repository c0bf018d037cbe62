"""The "tf32-class" precision mode (hriemo/precise.py) on a B200: north_star's bar is logits within 1e-4 of the
reference's fp32 forward.  Checked against the committed golden outputs of the unmodified reference and the float64
oracle; the mode's own kernels (split3, fp32 attention, masked mean, blend) against torch."""
import math

import pytest
import torch

import golden_util as G
import hriemo_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

LOGIT_TOL = 1e-4   # north_star, tf32
BETA_TOL = 1e-5
Z_TOL = 5e-4       # z is O(1..3) after LayerNorm


def _precise(model, ins, **kw):
    from hriemo import precise

    model = model.to(DEV)
    with precise.mode("tf32x3"):
        out = model(*[G.to_dev(x, DEV) for x in ins], **kw)
    torch.cuda.synchronize()
    assert precise.get_mode() == "bf16"
    return out


@pytest.mark.parametrize("name", ["cfg2_iemocap_nomask", "cfg2_iemocap_ragged", "ns_500x64_ragged",
                                  "utter_2d_inputs", "cfg3_mosei_default", "cfg3_mosei_v2"])
def test_tf32x3_forward_matches_reference_golden(name):
    fx = G.load(name)
    model, ins = G.build_fusion(fx)
    lo, be, z = [x.cpu() for x in _precise(model, ins)]
    assert lo.dtype == torch.float32 and lo.shape == fx["logits"].shape and z.shape == fx["z"].shape
    err = (lo - fx["logits"]).abs().max().item()
    assert err <= LOGIT_TOL, f"logits max-abs {err}"
    assert (be - fx["beta"]).abs().max().item() <= BETA_TOL
    assert (z - fx["z"]).abs().max().item() <= Z_TOL
    assert torch.equal(lo > 0, fx["logits"] > 0) and torch.equal(lo.argmax(-1), fx["logits"].argmax(-1))
    assert torch.equal(be > 0.5, fx["beta"] > 0.5)
    # and the default mode still answers within ITS bar on the same module (the two prepared operand sets coexist)
    lo16 = model(*[G.to_dev(x, DEV) for x in ins])[0].cpu()
    assert (lo16 - fx["logits"]).abs().max().item() <= 1e-2
    assert (lo16 - lo).abs().max().item() > 0.0


def test_tf32x3_attention_maps_against_oracle():
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    fx = G.load("tiny_explicit_weights")
    m = FusionWithEmotionDecoder(**fx["ctor"]).eval()
    m.load_state_dict(fx["state_dict"], strict=True)
    lo, be, z, pack = _precise(m, (fx["h_a"], fx["h_t"], fx["mask_a"], fx["mask_t"]), return_attention=True)
    sd = O.cast_state(fx["state_dict"], torch.float64)
    lo_o, be_o, z_o, pack_o = O.fusion_with_emotion_decoder(
        sd, fx["h_a"].double(), fx["h_t"].double(), fx["mask_a"], fx["mask_t"], n_heads=fx["ctor"]["n_heads"],
        return_attention=True)
    assert (lo.cpu() - lo_o).abs().max().item() <= LOGIT_TOL
    assert (be.cpu() - be_o).abs().max().item() <= BETA_TOL
    assert (z.cpu() - z_o).abs().max().item() <= Z_TOL
    for mine, ref in zip(pack["encoder"], pack_o["encoder"]):
        for k in ("audio_self", "text_self", "audio_queries_text", "text_queries_audio"):
            assert mine[k].shape == ref[k].shape and (mine[k].cpu() - ref[k]).abs().max().item() <= 1e-5, k
    for mine, ref in zip(pack["decoder"], pack_o["decoder"]):
        assert mine.shape == ref.shape and (mine.cpu() - ref).abs().max().item() <= 1e-5


def test_tf32x3_decisions_on_256_utterances():
    """Decision statistic of the mode against the float64 oracle on a d=192 model (256 utterances, ragged masks)."""
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(11)
    m = FusionWithEmotionDecoder(d_model=192, n_heads=2, beta_hidden=64, num_emotions=4).eval()
    ins = G.make_inputs(5, 256, 120, 32, 192, 192, True)
    sd = O.cast_state(m.state_dict(), torch.float64)
    lo_o, be_o, _ = O.fusion_with_emotion_decoder(sd, ins[0].double(), ins[1].double(), ins[2], ins[3], n_heads=2)
    lo, be, _ = [x.cpu() for x in _precise(m, ins)]
    assert (lo - lo_o).abs().max().item() <= LOGIT_TOL
    assert (be - be_o).abs().max().item() <= BETA_TOL
    clear = lo_o.abs() > LOGIT_TOL
    assert torch.equal((lo > 0)[clear], (lo_o > 0)[clear])
    assert torch.equal(lo.argmax(-1), lo_o.argmax(-1))


def test_tf32x3_slabs_are_independent(monkeypatch):
    from hriemo import precise
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(3)
    m = FusionWithEmotionDecoder(d_model=192, n_heads=2, beta_hidden=64).eval()
    ins = G.make_inputs(9, 7, 60, 20, 192, 192, True)
    whole = [x.cpu() for x in _precise(m, ins)]
    monkeypatch.setattr(precise, "MAX_ROWS_PER_SLAB", 120)   # two utterances per slab
    parts = [x.cpu() for x in _precise(m, ins)]
    for a, b in zip(whole, parts):
        assert torch.equal(a, b)


def test_precise_mode_rejects_unknown_names():
    from hriemo import lib as L, precise

    with pytest.raises(L.HriemoError):
        precise.set_mode("fp64")


# --------------------------------------------------------------------------- kernels of the mode
@pytest.mark.parametrize("K", [768, 74, 300])
def test_split3_gemm_reproduces_fp32_linear(K):
    from hriemo import lib as L, ops

    g = torch.Generator().manual_seed(K)
    x = torch.randn(257, K, generator=g) * 3.0
    w = torch.randn(96, K, generator=g) / math.sqrt(K)
    b = torch.randn(96, generator=g)
    x3, w3 = ops.split3(x.to(DEV)), ops.split3(w.to(DEV), weight=True)
    Kp = (K + 7) // 8 * 8
    assert x3.shape == (257, 3 * Kp) and w3.shape == (96, 3 * Kp)
    hi, lo = x3[:, :K].float().cpu(), x3[:, Kp:Kp + K].float().cpu()
    assert torch.equal(hi, x.bfloat16().float()) and torch.equal(x3[:, 2 * Kp:2 * Kp + K].float().cpu(), hi)
    assert (hi + lo - x).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()
    assert torch.equal(w3[:, Kp:Kp + K].float().cpu(), w.bfloat16().float())            # weights: [hi | hi | lo]
    if Kp > K:
        assert x3[:, K:Kp].abs().max().item() == 0.0
    y = ops.gemm(x3, w3, b.to(DEV), L.EPI_BIAS_F32).cpu()
    ref = x.double() @ w.double().t() + b.double()
    assert (y - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    r3 = ops.split3(x.to(DEV), relu=True)
    assert torch.equal(r3[:, :K].float().cpu(), x.clamp_min(0).bfloat16().float())


@pytest.mark.parametrize("B,H,Tq,Tk,dh,masked", [(3, 2, 37, 53, 96, True), (2, 4, 4, 64, 64, True), (2, 1, 20, 9, 32, False),
                                                  (1, 8, 130, 500, 96, True)])
def test_attention_f32_matches_torch(B, H, Tq, Tk, dh, masked):
    from hriemo import ops

    g = torch.Generator().manual_seed(B * 100 + Tk)
    d = H * dh
    qkv_q = torch.randn(B * Tq, d, generator=g)
    kv = torch.randn(B * Tk, 2 * d, generator=g)
    pad = O.ragged_masks(B, Tk, g) if masked else None
    out, probs = ops.attention_f32(qkv_q.to(DEV), kv[:, :d].to(DEV), kv.to(DEV)[:, d:], G.to_dev(pad, DEV), B, H, Tq, Tk, dh,
                                   want_probs=True)
    q = qkv_q.double().view(B, Tq, H, dh).transpose(1, 2)
    k = kv[:, :d].double().view(B, Tk, H, dh).transpose(1, 2)
    v = kv[:, d:].double().view(B, Tk, H, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(dh)
    if pad is not None:
        s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    ref = (p @ v).transpose(1, 2).reshape(B * Tq, d)
    assert (out.cpu() - ref).abs().max().item() <= 2e-5
    assert (probs.cpu() - p.mean(1)).abs().max().item() <= 1e-6


def test_masked_mean_and_blend_f32_match_torch():
    from hriemo import ops

    g = torch.Generator().manual_seed(4)
    B, T_a, L, d = 5, 33, 12, 192
    a, t = torch.randn(B, T_a, d, generator=g), torch.randn(B, L, d, generator=g)
    pad = O.ragged_masks(B, T_a, g)
    pad[2] = True   # a fully padded utterance pools to 0 (reference: clamp(min=1))
    pooled = ops.masked_mean_f32(a.view(B * T_a, d).to(DEV), pad.to(DEV), B, T_a).cpu()
    assert (pooled - O.masked_mean(a.double(), pad)).abs().max().item() <= 1e-6
    nomask = ops.masked_mean_f32(a.view(B * T_a, d).to(DEV), None, B, T_a).cpu()
    assert (nomask - a.double().mean(1)).abs().max().item() <= 1e-6
    w = torch.rand(B, d, generator=g)
    h, beta = ops.gate_blend_f32(a.view(B * T_a, d).to(DEV), T_a, t.view(B * L, d).to(DEV), w.to(DEV), B, L)
    ref = w[:, None] * a[:, :L] + (1 - w[:, None]) * t
    assert (h.cpu().view(B, L, d) - ref).abs().max().item() <= 1e-6
    assert (beta.cpu() - w.mean(-1, keepdim=True)).abs().max().item() <= 1e-6
