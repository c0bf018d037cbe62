"""Parity of the CUDA path (drop-in nn.Modules -> C ABI -> sm_100a kernels) on a B200.

Checked against (a) the committed golden outputs of the reference's fp32 forward and
(b) the float64 CPU oracle on the same seeded inputs.  Tolerances are north_star's bf16
bars: logits max-abs <= 1e-2, identical threshold decisions (logit > 0) on >= 99.9 % of
samples, identical label argmax, identical beta > 0.5 decision."""
import pytest
import torch

import golden_util as G
import hriemo_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

LOGIT_TOL = 1e-2     # north_star, bf16
BETA_TOL = 1e-4      # gate runs in fp32 on bf16-stored streams (measured ~5e-6)
Z_TOL = 6e-2         # z is O(1..3) after LayerNorm; bf16 GEMM operands (simulated: ~1e-2)


def _decisions(lo, ref):
    """threshold agreement excluding |ref logit| < tol, argmax agreement."""
    clear = ref.abs() > LOGIT_TOL
    thr = ((lo > 0) == (ref > 0))[clear].float().mean().item() if clear.any() else 1.0
    top2 = ref.topk(2, dim=-1).values
    unambiguous = (top2[:, 0] - top2[:, 1]) > 2 * LOGIT_TOL
    arg = (lo.argmax(-1) == ref.argmax(-1))[unambiguous].float().mean().item() if unambiguous.any() else 1.0
    return thr, arg


def _run(model, ins, **kw):
    model = model.to(DEV)
    out = model(*[G.to_dev(x, DEV) for x in ins], **kw)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", ["cfg2_iemocap_nomask", "cfg2_iemocap_ragged", "ns_500x64_ragged",
                                  "utter_2d_inputs", "cfg3_mosei_default", "cfg3_mosei_v2"])
def test_forward_matches_reference_golden(name):
    fx = G.load(name)
    model, ins = G.build_fusion(fx)
    lo, be, z = _run(model, ins)
    lo, be, z = lo.cpu(), be.cpu(), z.cpu()
    assert lo.dtype == torch.float32 and be.dtype == torch.float32
    assert lo.shape == fx["logits"].shape and be.shape == fx["beta"].shape and z.shape == fx["z"].shape
    err = (lo - fx["logits"]).abs().max().item()
    assert err <= LOGIT_TOL, f"logits max-abs {err}"
    assert (be - fx["beta"]).abs().max().item() <= BETA_TOL
    assert (z - fx["z"]).abs().max().item() <= Z_TOL
    thr, arg = _decisions(lo, fx["logits"])
    assert thr >= 0.999 and arg == 1.0
    assert torch.equal(be > 0.5, fx["beta"] > 0.5)


def test_forward_matches_oracle_with_attention_maps():
    """Tiny explicit-weight model: logits / beta / z and every head-averaged attention map."""
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    fx = G.load("tiny_explicit_weights")
    m = FusionWithEmotionDecoder(**fx["ctor"]).eval()
    m.load_state_dict(fx["state_dict"], strict=True)
    lo, be, z, pack = _run(m, (fx["h_a"], fx["h_t"], fx["mask_a"], fx["mask_t"]), return_attention=True)
    sd = O.cast_state(fx["state_dict"], torch.float64)
    lo_o, be_o, z_o, pack_o = O.fusion_with_emotion_decoder(
        sd, fx["h_a"].double(), fx["h_t"].double(), fx["mask_a"], fx["mask_t"], n_heads=fx["ctor"]["n_heads"],
        return_attention=True)
    assert (lo.cpu() - lo_o).abs().max().item() <= LOGIT_TOL
    assert (lo.cpu() - fx["logits"]).abs().max().item() <= LOGIT_TOL
    assert (be.cpu() - be_o).abs().max().item() <= BETA_TOL
    assert (z.cpu() - z_o).abs().max().item() <= Z_TOL
    assert set(pack) == {"encoder", "decoder"} and len(pack["encoder"]) == 2 and len(pack["decoder"]) == 2
    for mine, ref in zip(pack["encoder"], pack_o["encoder"]):
        for k in ("audio_self", "text_self", "audio_queries_text", "text_queries_audio"):
            assert mine[k].shape == ref[k].shape
            # probabilities in [0,1] from bf16 q/k: 2e-2 absolute
            assert (mine[k].cpu() - ref[k]).abs().max().item() <= 2e-2, k
            assert torch.allclose(mine[k].sum(-1).cpu(), torch.ones(mine[k].shape[:-1]), atol=1e-4)
    for mine, ref in zip(pack["decoder"], pack_o["decoder"]):
        assert mine.shape == ref.shape and (mine.cpu() - ref).abs().max().item() <= 2e-2


def test_forward_vs_oracle_medium_model_ragged():
    """d=192 / 2 heads (head dim 96, the IEMOCAP head dim) against the fp64 oracle."""
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    torch.manual_seed(7)
    m = FusionWithEmotionDecoder(d_model=192, n_heads=2, beta_hidden=64, num_emotions=5).eval()
    ins = G.make_inputs(8, 6, 150, 40, 192, 192, True)
    sd = O.cast_state(m.state_dict(), torch.float64)
    lo_o, be_o, z_o = O.fusion_with_emotion_decoder(sd, ins[0].double(), ins[1].double(), ins[2], ins[3], n_heads=2)
    lo, be, z = _run(m, ins)
    assert (lo.cpu() - lo_o).abs().max().item() <= LOGIT_TOL
    assert (be.cpu() - be_o).abs().max().item() <= BETA_TOL
    assert (z.cpu() - z_o).abs().max().item() <= Z_TOL


def test_fusion_classifier_config1():
    from models.fusion_classifier import FusionClassifier

    fx = G.load("cfg1_fusion_classifier")
    torch.manual_seed(fx["model_seed"])
    m = FusionClassifier().eval()
    G.assert_same_weights(m, fx["weights"])
    u = fx["utter"]
    g = torch.Generator().manual_seed(u["in_seed"])
    h_a, h_t = torch.randn(u["B"], 768, generator=g), torch.randn(u["B"], 768, generator=g)
    lo, be, pooled = _run(m, (h_a, h_t))
    assert lo.shape == (32, 4) and be.shape == (32, 1) and pooled.shape == (32, 768)
    assert (lo.cpu() - u["logits"]).abs().max().item() <= LOGIT_TOL
    assert (be.cpu() - u["beta"]).abs().max().item() <= BETA_TOL
    assert (pooled.cpu() - u["pooled"]).abs().max().item() <= Z_TOL
    assert torch.equal(lo.cpu().argmax(-1), u["logits"].argmax(-1)) or _decisions(lo.cpu(), u["logits"])[1] == 1.0
    s = fx["seq"]
    ins = G.make_inputs(s["in_seed"], s["B"], s["T_a"], s["T_t"], 768, 768, True)
    lo, be, pooled = _run(m, ins)
    assert (lo.cpu() - s["logits"]).abs().max().item() <= LOGIT_TOL
    assert (be.cpu() - s["beta"]).abs().max().item() <= BETA_TOL
    assert (pooled.cpu() - s["pooled"]).abs().max().item() <= Z_TOL


def test_legacy_block_and_scalar_gate():
    """The reference's own test scenarios (tests/test_beta_gate.py, tests/test_cross_modal_block.py)."""
    from models.beta_gate import BetaGate
    from models.cross_modal_block import CrossModalTransformer

    fx = G.load("legacy_block_scalar_gate")
    torch.manual_seed(fx["model_seed"])
    cross = CrossModalTransformer(num_layers=2, d_model=768, n_heads=8).eval().to(DEV)
    gate = BetaGate(d_model=768, hidden_dim=256).eval().to(DEV)
    G.assert_same_weights(cross, fx["weights_cross"])
    u = fx["utter"]
    g = torch.Generator().manual_seed(u["in_seed"])
    h_a, h_t = torch.randn(32, 1, 768, generator=g).to(DEV), torch.randn(32, 1, 768, generator=g).to(DEV)
    a, t = cross(h_a, h_t)
    hf, beta = gate(a, t)
    assert hf.shape == (32, 1, 768) and beta.shape == (32, 1)
    assert (a.cpu() - u["h_a_tilde"]).abs().max().item() <= Z_TOL
    assert (t.cpu() - u["h_t_tilde"]).abs().max().item() <= Z_TOL
    assert (hf.cpu() - u["h_fusion"]).abs().max().item() <= Z_TOL
    assert (beta.cpu() - u["beta"]).abs().max().item() <= 2e-3  # scalar gate sees bf16-rounded streams
    s = fx["seq"]
    g = torch.Generator().manual_seed(s["in_seed"])
    s_a, s_t = torch.randn(8, 400, 768, generator=g).to(DEV), torch.randn(8, 128, 768, generator=g).to(DEV)
    zm_a = torch.zeros(8, 400, dtype=torch.bool, device=DEV)
    zm_t = torch.zeros(8, 128, dtype=torch.bool, device=DEV)
    a, t = cross(s_a, s_t, zm_a, zm_t)
    hf, beta = gate(a, t, zm_a, zm_t)
    assert a.shape == (8, 400, 768) and t.shape == (8, 128, 768)
    assert (a.cpu()[:, ::40, ::32] - s["h_a_tilde_slice"]).abs().max().item() <= Z_TOL
    assert (t.cpu()[:, ::16, ::32] - s["h_t_tilde_slice"]).abs().max().item() <= Z_TOL
    assert (hf.cpu()[:, ::16, ::32] - s["h_fusion_slice"]).abs().max().item() <= Z_TOL
    assert (beta.cpu() - s["beta"]).abs().max().item() <= 2e-3


def test_size_independent_properties_at_scale():
    """Properties that need no oracle, at a batch the oracle could not finish in seconds:
    utterances are independent (slab / shard invariance, permutation equivariance), PAD key
    content is irrelevant, masks of all-False equal no mask, outputs are deterministic."""
    from models import fusion_with_emotion_decoder as F

    torch.manual_seed(1234)
    m = F.FusionWithEmotionDecoder().eval().to(DEV)
    B, T_a, T_t = 96, 300, 50
    h_a, h_t, m_a, m_t = [G.to_dev(x, DEV) for x in G.make_inputs(77, B, T_a, T_t, 768, 768, True)]
    lo, be, z = m(h_a, h_t, m_a, m_t)
    lo2, be2, z2 = m(h_a, h_t, m_a, m_t)
    assert torch.equal(lo, lo2) and torch.equal(be, be2) and torch.equal(z, z2)          # deterministic
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(DEV)
    lo_p, be_p, _ = m(h_a[perm], h_t[perm], m_a[perm], m_t[perm])
    assert torch.equal(lo_p, lo[perm]) and torch.equal(be_p, be[perm])                   # permutation equivariance
    lo_s, be_s, _ = m(h_a[:17], h_t[:17], m_a[:17], m_t[:17])
    assert torch.equal(lo_s, lo[:17]) and torch.equal(be_s, be[:17])                     # shard invariance
    old = F.MAX_ROWS_PER_SLAB
    try:
        F.MAX_ROWS_PER_SLAB = 7 * T_a                                                    # force ragged slabs
        lo_c, be_c, z_c = m(h_a, h_t, m_a, m_t)
    finally:
        F.MAX_ROWS_PER_SLAB = old
    assert torch.equal(lo_c, lo) and torch.equal(be_c, be) and torch.equal(z_c, z)
    junk_a = torch.where(m_a[..., None], torch.full_like(h_a, 9.0), h_a)                 # PAD content irrelevant
    junk_t = torch.where(m_t[..., None], torch.full_like(h_t, -9.0), h_t)
    lo_j, be_j, _ = m(junk_a, junk_t, m_a, m_t)
    assert torch.equal(lo_j, lo) and torch.equal(be_j, be)
    f_a, f_t = torch.zeros_like(m_a), torch.zeros_like(m_t)
    lo_n, be_n, _ = m(h_a, h_t)
    lo_f, be_f, _ = m(h_a, h_t, f_a, f_t)
    assert torch.equal(lo_n, lo_f) and torch.equal(be_n, be_f)                           # all-False mask == None
    assert torch.isfinite(lo).all() and ((be > 0) & (be < 1)).all()


def test_nan_for_fully_padded_sample_and_errors():
    from hriemo.lib import HriemoError
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    fx = G.load("tiny_explicit_weights")
    m = FusionWithEmotionDecoder(**fx["ctor"]).eval()
    m.load_state_dict(fx["state_dict"])
    m_t = fx["mask_t"].clone()
    m_t[1] = True
    lo, _, _ = _run(m, (fx["h_a"], fx["h_t"], fx["mask_a"], m_t))
    assert torch.isnan(lo[1]).all() and torch.isfinite(lo[0]).all() and torch.isfinite(lo[2]).all()
    with pytest.raises(ValueError, match="Expected 2D or 3D tensor"):
        m(fx["h_a"].to(DEV)[None], fx["h_t"].to(DEV))
    with pytest.raises(HriemoError, match="no CPU fallback"):
        m(fx["h_a"], fx["h_t"].to(DEV))
    with pytest.raises(RuntimeError):
        m(fx["h_a"].to(DEV), fx["h_t"].to(DEV), fx["mask_t"].to(DEV), None)  # wrong mask length


def test_load_state_dict_refreshes_prepared_weights():
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    fx = G.load("tiny_explicit_weights")
    m = FusionWithEmotionDecoder(**fx["ctor"]).eval().to(DEV)
    ins = [G.to_dev(fx[k], DEV) for k in ("h_a", "h_t", "mask_a", "mask_t")]
    lo_rand, _, _ = m(*ins)
    m.load_state_dict(fx["state_dict"], strict=True)  # in-place copy_: bf16 operand cache must be rebuilt
    lo, _, _ = m(*ins)
    assert (lo.cpu() - fx["logits"]).abs().max().item() <= LOGIT_TOL
    assert (lo_rand.cpu() - fx["logits"]).abs().max().item() > LOGIT_TOL


@pytest.mark.parametrize("name,slab", [("cfg2_iemocap_ragged", 3), ("cfg3_mosei_default", 2), ("cfg2_iemocap_nomask", 64)])
def test_forward_from_host_equals_device_forward(name, slab):
    """The end-to-end entry (pinned host features, staged slab by slab on a side stream) returns
    exactly what the device-resident forward returns: slabs are independent utterances."""
    from hriemo import pipeline

    fx = G.load(name)
    model, ins = G.build_fusion(fx)
    model = model.to(DEV)
    host = [None if x is None else x.clone().pin_memory() for x in (list(ins) + [None, None])[:4]]
    ref = model(*[G.to_dev(x, DEV) for x in ins])
    torch.cuda.synchronize()
    # default plan = "auto" (copy engine from the front, host cast from the back: pipeline.TwoEndedPlan) when the
    # batch has three or more slabs; the fixed plans with every slab / no slab pre-cast by the host; one thread
    for kw in ({}, {}, dict(host_cast_every=1), dict(host_cast_every=0), dict(host_cast_every="auto")):
        pipeline.reset_stats()
        lo, be, z = pipeline.forward_from_host(model, *host, device=DEV, slab=slab, **kw)   # repeated calls reuse the staging sets
        assert lo.device.type == "cpu"
        # same kernels on the same rows; only the slab boundaries (and who cast a slab to bf16) differ
        assert torch.equal(lo, ref[0].cpu()) and torch.equal(be, ref[1].cpu()) and torch.equal(z, ref[2].cpu()), kw
        B = host[0].shape[0]
        assert pipeline.STATS["slabs"] == (B + min(slab, B) - 1) // min(slab, B) and pipeline.STATS["h2d_bytes"] > 0


def test_full_size_properties_north_star():
    """BASELINE.json's full size (B = 4096, T_a = 500, T_t = 64, default model), where the oracle is too
    slow for every sample: size-independent properties of the path plus an oracle spot check.
      (1) utterances are independent: permuting the batch permutes the outputs bit for bit
          (this crosses slab boundaries: B*T_a = 2.05 M rows = two slabs);
      (2) a PAD tail behind a mask changes nothing that is not masked: extending T_a by 12 PAD frames
          leaves logits / beta within the parity tolerance;
      (3) eight samples drawn from the batch agree with the float64 CPU oracle."""
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    B, T_a, T_t = 4096, 500, 64
    torch.manual_seed(1234)
    m = FusionWithEmotionDecoder().eval()
    sd = O.cast_state(m.state_dict(), torch.float64)
    m = m.to(DEV)
    g = torch.Generator(device=DEV).manual_seed(4321)
    h_a = torch.randn(B, T_a, 768, generator=g, device=DEV)
    h_t = torch.randn(B, T_t, 768, generator=g, device=DEV)
    gc = torch.Generator().manual_seed(99)
    m_a, m_t = O.ragged_masks(B, T_a, gc).to(DEV), O.ragged_masks(B, T_t, gc).to(DEV)
    lo, be, z = m(h_a, h_t, m_a, m_t)
    assert torch.isfinite(lo).all() and torch.isfinite(be).all() and torch.isfinite(z).all()

    perm = torch.randperm(B, generator=gc).to(DEV)
    lo_p, be_p, z_p = m(h_a[perm], h_t[perm], m_a[perm], m_t[perm])
    assert torch.equal(lo_p, lo[perm]) and torch.equal(be_p, be[perm]) and torch.equal(z_p, z[perm])
    del lo_p, be_p, z_p

    sub = slice(0, 512)  # (2) on one slab's worth keeps the test's memory modest
    pad = 12
    h_a2 = torch.cat([h_a[sub], torch.randn(512, pad, 768, generator=g, device=DEV) * 50.0], dim=1)
    m_a2 = torch.cat([m_a[sub], torch.ones(512, pad, dtype=torch.bool, device=DEV)], dim=1)
    lo2, be2, _ = m(h_a2, h_t[sub], m_a2, m_t[sub])
    assert (lo2 - lo[sub]).abs().max().item() <= LOGIT_TOL and (be2 - be[sub]).abs().max().item() <= BETA_TOL

    idx = torch.tensor([0, 1, 777, 2047, 2048, 3000, 4094, 4095])
    lo_o, be_o, z_o = O.fusion_with_emotion_decoder(sd, h_a[idx].double().cpu(), h_t[idx].double().cpu(),
                                                    m_a[idx].cpu(), m_t[idx].cpu(), n_heads=8)
    assert (lo[idx].cpu() - lo_o).abs().max().item() <= LOGIT_TOL
    assert (be[idx].cpu() - be_o).abs().max().item() <= BETA_TOL
    assert (z[idx].cpu() - z_o).abs().max().item() <= Z_TOL
    assert torch.equal(be[idx].cpu() > 0.5, be_o > 0.5)


def test_model_on_second_gpu_with_other_current_device():
    """The C ABI launches on the current device: the wrappers must switch to the tensors' device
    (a model on cuda:1 while cuda:0 is current), including the per-device shared-memory opt-in."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    fx = G.load("cfg2_iemocap_ragged")
    model, ins = G.build_fusion(fx)
    model = model.to("cuda:1")
    assert torch.cuda.current_device() == 0
    lo, be, z = model(*[G.to_dev(x, "cuda:1") for x in ins])
    torch.cuda.synchronize("cuda:1")
    assert lo.device.index == 1
    assert (lo.cpu() - fx["logits"]).abs().max().item() <= LOGIT_TOL
    assert (be.cpu() - fx["beta"]).abs().max().item() <= BETA_TOL


def test_run_split_writes_the_reference_inference_outputs(tmp_path):
    """hriemo.infer.run_split: the files scripts/infer/mosei_eval_infer.py:237-284 writes, same names, dtypes and
    nesting; probabilities against the reference's golden logits."""
    import numpy as np
    from hriemo import infer

    fx = G.load("cfg3_mosei_default")
    model, (h_a, h_t, m_a, m_t) = G.build_fusion(fx)
    model = model.to(DEV)
    B = h_a.shape[0]
    y = (torch.arange(B * 6).view(B, 6) % 2).float()
    batches = [(h_a[:3], m_a[:3], h_t[:3], m_t[:3], y[:3]), (h_a[3:], m_a[3:], h_t[3:], m_t[3:], y[3:])]
    thr = [0.5, 0.4, 0.6, 0.5, 0.3, 0.7]
    out = infer.run_split(model, batches, DEV, str(tmp_path), "val", dump_beta=True, dump_attn=True, attn_max_samples=3,
                          thresholds=thr)
    prob = np.load(tmp_path / "val_y_prob.npy")
    assert prob.dtype == np.float32 and prob.shape == (B, 6)
    assert np.abs(prob - torch.sigmoid(fx["logits"]).numpy()).max() <= 2.5e-3      # d sigmoid <= 1/4 of the 1e-2 logit bar
    assert np.array_equal(np.load(tmp_path / "val_y_true.npy"), y.numpy())
    beta = np.load(tmp_path / "val_beta_mean.npy")
    assert beta.shape == (B,) and np.abs(beta - fx["beta"].numpy()[:, 0]).max() <= BETA_TOL
    pred = np.load(tmp_path / "val_y_pred.npy")
    assert pred.dtype == np.uint8 and np.array_equal(pred, (prob >= np.asarray(thr, dtype=np.float32)[None]).astype(np.uint8))
    att = torch.load(tmp_path / "val_attentions.pt", weights_only=False)
    assert set(att) == {"encoder", "decoder"} and len(att["encoder"]) == 1 and out["attn_samples"] == 3   # capped after batch 1
    layers = att["encoder"][0]
    assert len(layers) == 2 and set(layers[0]) == {"audio_self", "text_self", "audio_queries_text", "text_queries_audio"}
    assert layers[0]["audio_self"].shape == (3, h_a.shape[1], h_a.shape[1]) and isinstance(layers[0]["audio_self"], np.ndarray)
    assert att["decoder"][0][0].shape == (3, 6, h_t.shape[1])


def test_forward_from_host_two_calls_in_flight():
    """wait=False: a second call is issued before the first is awaited (its copies overlap the first call's
    compute and both share the persistent staging sets).  Different inputs per call: each call must return
    exactly the device-resident forward of ITS batch, for dense and bucketed plans."""
    from hriemo import pipeline

    fx = G.load("cfg2_iemocap_ragged")
    model, ins = G.build_fusion(fx)
    model = model.to(DEV)
    g = torch.Generator().manual_seed(77)
    batches = []
    for k in range(4):
        B = 11
        h_a, h_t = torch.randn(B, 300, 768, generator=g), torch.randn(B, 50, 768, generator=g)
        m_a, m_t = O.ragged_masks(B, 300, g), O.ragged_masks(B, 50, g)
        batches.append([x.pin_memory() for x in (h_a, h_t, m_a, m_t)])
    want = [[o.cpu() for o in model(*[x.to(DEV) for x in b])] for b in batches]
    torch.cuda.synchronize()
    for kw in (dict(slab=2, host_cast_every=2), dict(slab=3, host_cast_every=1), dict(slab=4, bucket=True), dict(slab=2)):
        pend, got = None, []
        for b in batches:
            nxt = pipeline.forward_from_host(model, *b, device=DEV, wait=False, **kw)
            assert isinstance(nxt, pipeline.PendingResult)
            if pend is not None:
                got.append(pend.wait())
            pend = nxt
        got.append(pend.wait())
        for res, ref in zip(got, want):
            if kw.get("bucket"):
                assert (res[0] - ref[0]).abs().max().item() <= 2e-3 and (res[1] - ref[1]).abs().max().item() <= 1e-5
            else:
                assert torch.equal(res[0], ref[0]) and torch.equal(res[1], ref[1]) and torch.equal(res[2], ref[2])


@pytest.mark.parametrize("workload", ["cfg2", "cfg3"])
def test_decision_statistics_over_1024_utterances(workload):
    """north_star: "identical per-emotion threshold decisions on >= 99.9 % of samples" needs a denominator that can
    show 99.9 %: 1024 utterances (half with ragged masks) of BASELINE configs 2 (IEMOCAP 300/50) and 3 (MOSEI
    wrapper), the oracle port on the host cores against the GPU forward on the same weights and inputs -- the very
    leg bench.py reports as `parity` (bench.parity_and_cpu_leg), so the driver-run artefact and this test agree."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_parity", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = dict(bench.WORKLOADS[workload])
    wl["B"] = 8                                     # the resident batch is not used by this leg
    model, _, _, d_a, d_t, n_heads, mosei = bench.build_model_and_inputs(wl, torch.device(DEV), 0)
    parity, cpu = bench.parity_and_cpu_leg(model, torch.device(DEV), wl["T_a"], wl["T_t"], d_a, d_t, n_heads, mosei, 1024, 128)
    print(parity)
    n_e = 6 if mosei else 4
    assert parity["n"] == 1024 and parity["n_ragged"] == 512 and parity["decisions"] == 1024 * n_e
    assert parity["logits_max_abs"] <= LOGIT_TOL, parity
    assert parity["beta_max_abs"] <= 1e-4, parity
    assert parity["thr_agree"] >= 0.999 and parity["thr_disagreements_outside_tol"] <= 1024 * n_e // 1000, parity
    assert parity["argmax_agree"] >= 0.999, parity
    assert parity["beta_gt_half_agree"] == 1.0, parity
    assert parity["thr_excluded"] < parity["decisions"] // 2, parity      # the statistic rests on a real denominator
