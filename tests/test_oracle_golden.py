"""Pins the CPU oracle (oracle/hriemo_oracle.py) against outputs of the reference itself
(tests/golden/*.pt, produced by tests/golden/make_golden.py from /root/reference).

The oracle runs in float64; the fixtures hold the reference's float32 outputs, so the
agreement bound is float32 round-off of a ~50-op deep network: 2e-5 absolute."""
import pytest
import torch

import golden_util as G
import hriemo_oracle as O

TOL = 2e-5


def _f64(x):
    return None if x is None else (x.double() if x.is_floating_point() else x)


@pytest.mark.parametrize("name", ["cfg2_iemocap_nomask", "cfg2_iemocap_ragged", "ns_500x64_ragged",
                                  "utter_2d_inputs", "cfg3_mosei_default", "cfg3_mosei_v2"])
def test_oracle_matches_reference_outputs(name):
    fx = G.load(name)
    model, (h_a, h_t, m_a, m_t) = G.build_fusion(fx)
    sd = O.cast_state(model.state_dict(), torch.float64)
    H = G.n_heads_of(fx)
    if fx["kind"] == "mosei":
        lo, be, z = O.mosei_fusion_with_emotion_decoder(sd, _f64(h_a), _f64(h_t), m_a, m_t, n_heads=H)
    else:
        lo, be, z = O.fusion_with_emotion_decoder(sd, _f64(h_a), _f64(h_t), m_a, m_t, n_heads=H)
    assert (lo - fx["logits"]).abs().max().item() < TOL
    assert (be - fx["beta"]).abs().max().item() < TOL
    assert (z - fx["z"]).abs().max().item() < 5 * TOL  # z is O(1..3), logits O(0.5)
    assert lo.shape == fx["logits"].shape and be.shape == fx["beta"].shape and z.shape == fx["z"].shape


def test_oracle_tiny_explicit_weights_and_attention_maps():
    fx = G.load("tiny_explicit_weights")
    sd = O.cast_state(fx["state_dict"], torch.float64)
    lo, be, z, pack = O.fusion_with_emotion_decoder(sd, _f64(fx["h_a"]), _f64(fx["h_t"]), fx["mask_a"], fx["mask_t"],
                                                    n_heads=fx["ctor"]["n_heads"], return_attention=True)
    assert (lo - fx["logits"]).abs().max().item() < TOL
    assert (be - fx["beta"]).abs().max().item() < TOL
    assert (z - fx["z"]).abs().max().item() < 5 * TOL
    ref = fx["attn"]
    assert len(pack["encoder"]) == len(ref["encoder"]) == 2 and len(pack["decoder"]) == len(ref["decoder"]) == 2
    for mine, theirs in zip(pack["encoder"], ref["encoder"]):
        assert set(mine) == set(theirs) == {"audio_self", "text_self", "audio_queries_text", "text_queries_audio"}
        for k in mine:
            assert mine[k].shape == theirs[k].shape
            assert (mine[k] - theirs[k]).abs().max().item() < TOL
    for mine, theirs in zip(pack["decoder"], ref["decoder"]):
        assert mine.shape == theirs.shape and (mine - theirs).abs().max().item() < TOL


def test_oracle_fusion_classifier_config1():
    from models.fusion_classifier import FusionClassifier

    fx = G.load("cfg1_fusion_classifier")
    torch.manual_seed(fx["model_seed"])
    m = FusionClassifier().eval()
    G.assert_same_weights(m, fx["weights"])
    sd = O.cast_state(m.state_dict(), torch.float64)
    u = fx["utter"]
    g = torch.Generator().manual_seed(u["in_seed"])
    h_a, h_t = torch.randn(u["B"], 768, generator=g), torch.randn(u["B"], 768, generator=g)
    lo, be, pooled = O.fusion_classifier(sd, h_a.double(), h_t.double())
    assert (lo - u["logits"]).abs().max().item() < TOL
    assert (be - u["beta"]).abs().max().item() < TOL
    assert (pooled - u["pooled"]).abs().max().item() < 5 * TOL
    s = fx["seq"]
    h_a, h_t, m_a, m_t = G.make_inputs(s["in_seed"], s["B"], s["T_a"], s["T_t"], 768, 768, True)
    lo, be, pooled = O.fusion_classifier(sd, h_a.double(), h_t.double(), m_a, m_t)
    assert (lo - s["logits"]).abs().max().item() < TOL
    assert (be - s["beta"]).abs().max().item() < TOL
    assert (pooled - s["pooled"]).abs().max().item() < 5 * TOL


def test_oracle_legacy_block_and_scalar_gate():
    from models.beta_gate import BetaGate
    from models.cross_modal_block import CrossModalTransformer

    fx = G.load("legacy_block_scalar_gate")
    torch.manual_seed(fx["model_seed"])
    cross = CrossModalTransformer(num_layers=2, d_model=768, n_heads=8).eval()
    gate = BetaGate(d_model=768, hidden_dim=256).eval()
    G.assert_same_weights(cross, fx["weights_cross"])
    G.assert_same_weights(gate, fx["weights_gate"])
    sdc, sdg = O.cast_state(cross.state_dict(), torch.float64), O.cast_state(gate.state_dict(), torch.float64)
    u = fx["utter"]
    g = torch.Generator().manual_seed(u["in_seed"])
    h_a, h_t = torch.randn(u["B"], 1, 768, generator=g).double(), torch.randn(u["B"], 1, 768, generator=g).double()
    a, t, _ = O.cross_modal_transformer(sdc, "", h_a, h_t, None, None, 8, legacy=True)
    hf, beta = O.beta_gate_legacy(sdg, "", a, t, None, None)
    assert hf.shape == (32, 1, 768) and beta.shape == (32, 1)  # the shapes tests/test_beta_gate.py:25 states
    assert (a - u["h_a_tilde"]).abs().max().item() < 5 * TOL
    assert (t - u["h_t_tilde"]).abs().max().item() < 5 * TOL
    assert (hf - u["h_fusion"]).abs().max().item() < 5 * TOL
    assert (beta - u["beta"]).abs().max().item() < TOL
    s = fx["seq"]
    g = torch.Generator().manual_seed(s["in_seed"])
    s_a, s_t = torch.randn(8, 400, 768, generator=g).double(), torch.randn(8, 128, 768, generator=g).double()
    zm_a, zm_t = torch.zeros(8, 400, dtype=torch.bool), torch.zeros(8, 128, dtype=torch.bool)
    a, t, _ = O.cross_modal_transformer(sdc, "", s_a, s_t, zm_a, zm_t, 8, legacy=True)
    hf, beta = O.beta_gate_legacy(sdg, "", a, t, zm_a, zm_t)
    assert a.shape == (8, 400, 768) and t.shape == (8, 128, 768)  # tests/test_cross_modal_block.py
    assert (a[:, ::40, ::32] - s["h_a_tilde_slice"]).abs().max().item() < 5 * TOL
    assert (t[:, ::16, ::32] - s["h_t_tilde_slice"]).abs().max().item() < 5 * TOL
    assert (hf[:, ::16, ::32] - s["h_fusion_slice"]).abs().max().item() < 5 * TOL
    assert (beta - s["beta"]).abs().max().item() < TOL


def test_oracle_edge_semantics():
    """Semantics the reference exhibits but never tests."""
    fx = G.load("tiny_explicit_weights")
    sd = O.cast_state(fx["state_dict"], torch.float64)
    H = fx["ctor"]["n_heads"]
    h_a, h_t = fx["h_a"].double(), fx["h_t"].double()
    # (1) values at PAD key positions do not influence valid outputs
    m_a, m_t = fx["mask_a"], fx["mask_t"]
    lo0, be0, _ = O.fusion_with_emotion_decoder(sd, h_a, h_t, m_a, m_t, n_heads=H)
    h_a2 = torch.where(m_a[..., None], torch.full_like(h_a, 7.0), h_a)
    lo1, be1, _ = O.fusion_with_emotion_decoder(sd, h_a2, h_t, m_a, m_t, n_heads=H)
    # audio PAD rows only reach the output through padded *query* rows, which masked_mean and the
    # fused mask exclude — except the first T_t fused positions, which the fused mask also hides.
    assert (lo0 - lo1).abs().max().item() < 1e-9 and (be0 - be1).abs().max().item() < 1e-9
    # (2) a sample whose keys are all PAD yields NaN logits (torch.softmax over all -inf)
    m_t_bad = m_t.clone()
    m_t_bad[1] = True
    lo2, _, _ = O.fusion_with_emotion_decoder(sd, h_a, h_t, m_a, m_t_bad, n_heads=H)
    assert torch.isnan(lo2[1]).all() and torch.isfinite(lo2[0]).all()
    # (3) wrong rank raises ValueError like _ensure_3d (fusion_with_emotion_decoder.py:69)
    with pytest.raises(ValueError):
        O.fusion_with_emotion_decoder(sd, h_a[None], h_t, n_heads=H)
    # (4) utterances are independent: batch order / composition does not matter
    perm = torch.tensor([2, 0, 1])
    lo3, be3, _ = O.fusion_with_emotion_decoder(sd, h_a[perm], h_t[perm], m_a[perm], m_t[perm], n_heads=H)
    assert (lo3 - lo0[perm]).abs().max().item() < 1e-12
