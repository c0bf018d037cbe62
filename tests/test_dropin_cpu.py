"""CPU checks of the drop-in boundary: module paths, constructor
signatures, state_dict contract, seeded-init equality with the reference (when the
reference checkout is present), error behaviour without a GPU."""
import importlib
import inspect
import os
import sys
import types

import pytest
import torch

import golden_util as G

REF_DIR = "/root/reference/models"


def test_module_paths_and_constructor_signatures():
    exp = {
        ("models.cross_modal_block", "CrossModalBlock"): dict(d_model=768, n_heads=8, dropout=0.1),
        ("models.cross_modal_block", "CrossModalTransformer"): dict(num_layers=2, d_model=768, n_heads=8, dropout=0.1),
        ("models.cross_modal_block_tacfn", "CrossModalBlock"): dict(d_model=768, n_heads=8, dropout=0.1),
        ("models.cross_modal_block_tacfn", "CrossModalTransformer"): dict(num_layers=2, d_model=768, n_heads=8, dropout=0.1),
        ("models.beta_gate", "BetaGate"): dict(d_model=768, hidden_dim=256),
        ("models.beta_gate_tacfn", "BetaGate"): dict(d_model=768, hidden_dim=256),
        ("models.emotion_decoder", "EmotionDecoder"): dict(d_model=768, num_emotions=4, n_heads=8, num_layers=2,
                                                           dim_feedforward=2048, dropout=0.1, use_output_layer=True),
        ("models.fusion_with_emotion_decoder", "FusionWithEmotionDecoder"): dict(
            d_model=768, num_emotions=4, n_heads=8, num_layers_fusion=2, num_layers_decoder=2, beta_hidden=256, dropout=0.1),
        ("models.mosei_fusion_with_emotion_decoder", "MoseiFusionWithEmotionDecoder"): dict(
            d_model=256, num_emotions=6, n_heads=4, num_layers_fusion=2, num_layers_decoder=2, beta_hidden=128, dropout=0.2),
        ("models.fusion_classifier", "FusionClassifier"): dict(d_model=768, num_classes=4, n_heads=8, num_layers=2,
                                                               beta_hidden=256, dropout=0.2),
    }
    for (mod, cls), defaults in exp.items():
        c = getattr(importlib.import_module(mod), cls)
        sig = inspect.signature(c.__init__)
        got = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect._empty}
        assert got == defaults, (mod, cls, got)
    from models.beta_gate import masked_mean as mm1
    from models.beta_gate_tacfn import masked_mean as mm2
    assert callable(mm1) and callable(mm2)
    fwd = inspect.signature(importlib.import_module("models.fusion_with_emotion_decoder").FusionWithEmotionDecoder.forward)
    assert list(fwd.parameters) == ["self", "h_a", "h_t", "mask_a", "mask_t", "return_attention"]


def test_state_dict_contract_default_model():
    """119 tensors, 54 553 857 parameters, reference names and shapes."""
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    sd = FusionWithEmotionDecoder().state_dict()
    assert len(sd) == 119 and sum(v.numel() for v in sd.values()) == 54553857
    d = 768
    for l in range(2):
        p = f"cross_modal.layers.{l}."
        for a in ("self_attn_a", "self_attn_t", "attn_a2t", "attn_t2a"):
            assert sd[p + a + ".in_proj_weight"].shape == (3 * d, d) and sd[p + a + ".in_proj_bias"].shape == (3 * d,)
            assert sd[p + a + ".out_proj.weight"].shape == (d, d) and sd[p + a + ".out_proj.bias"].shape == (d,)
        for n in ("self_norm_a", "self_norm_t", "norm_a1", "norm_a2", "norm_t1", "norm_t2"):
            assert sd[p + n + ".weight"].shape == (d,)
        for f in ("ffn_a", "ffn_t"):
            assert sd[p + f + ".0.weight"].shape == (4 * d, d) and sd[p + f + ".2.weight"].shape == (d, 4 * d)
        q = f"emotion_decoder.layers.{l}."
        assert sd[q + "linear1.weight"].shape == (2048, d) and sd[q + "linear2.weight"].shape == (d, 2048)
        assert sd[q + "cross_attn.in_proj_weight"].shape == (3 * d, d)
    assert sd["beta_gate.mlp.0.weight"].shape == (256, 4 * d) and sd["beta_gate.mlp.2.weight"].shape == (d, 256)
    assert sd["emotion_decoder.emotion_queries"].shape == (4, d)
    assert sd["emotion_decoder.out_proj.weight"].shape == (1, d) and sd["emotion_decoder.out_proj.bias"].shape == (1,)


def test_other_state_dicts():
    from models.beta_gate import BetaGate
    from models.fusion_classifier import FusionClassifier
    from models.mosei_fusion_with_emotion_decoder import MoseiFusionWithEmotionDecoder

    sd = MoseiFusionWithEmotionDecoder(74, 300).state_dict()
    assert sum(v.numel() for v in sd.values()) == 7634561
    assert sd["audio_proj.weight"].shape == (256, 74) and sd["text_proj.weight"].shape == (256, 300)
    assert all(k.startswith(("audio_proj.", "text_proj.", "backbone.")) for k in sd)
    sd = FusionClassifier().state_dict()
    assert sum(v.numel() for v in sd.values()) == 39389444
    assert {"classifier.0.weight", "classifier.1.weight", "classifier.4.weight"} <= set(sd)
    sd = BetaGate().state_dict()
    assert sd["mlp.2.weight"].shape == (1, 256) and len(sd) == 4


def test_golden_weight_checksums_reproduced_from_seed():
    for name in ("cfg2_iemocap_ragged", "cfg3_mosei_default", "cfg3_mosei_v2", "utter_2d_inputs"):
        G.build_fusion(G.load(name))  # asserts the checksum


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference checkout not present (GPU box)")
def test_seeded_init_and_keys_identical_to_reference():
    pkg = types.ModuleType("refmodels")
    pkg.__path__ = [REF_DIR]
    sys.modules["refmodels"] = pkg
    pairs = [
        ("fusion_with_emotion_decoder", "FusionWithEmotionDecoder", (), {}),
        ("mosei_fusion_with_emotion_decoder", "MoseiFusionWithEmotionDecoder", (74, 300), {}),
        ("fusion_classifier", "FusionClassifier", (), {}),
        ("cross_modal_block", "CrossModalTransformer", (), {}),
        ("cross_modal_block_tacfn", "CrossModalTransformer", (), dict(num_layers=1, d_model=64, n_heads=2)),
        ("beta_gate", "BetaGate", (), {}),
        ("beta_gate_tacfn", "BetaGate", (), {}),
        ("emotion_decoder", "EmotionDecoder", (), dict(d_model=64, num_emotions=6, n_heads=2, use_output_layer=False)),
    ]
    for mod, cls, args, kw in pairs:
        torch.manual_seed(4321)
        mine = getattr(importlib.import_module("models." + mod), cls)(*args, **kw)
        torch.manual_seed(4321)
        ref = getattr(importlib.import_module("refmodels." + mod), cls)(*args, **kw)
        a, b = mine.state_dict(), ref.state_dict()
        assert list(a) == list(b), (mod, cls)
        for k in a:
            assert torch.equal(a[k], b[k]), (mod, cls, k)
        # strict load both ways (scripts/infer/mosei_eval_infer.py:340-341)
        mine.load_state_dict(ref.state_dict(), strict=True)
        ref.load_state_dict(mine.state_dict(), strict=True)


def test_nn_module_behaviour():
    """Plain nn.Module: parameters for an optimizer, attribute attach, train/eval, .to()."""
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    m = FusionWithEmotionDecoder(d_model=64, n_heads=2, beta_hidden=16)
    assert all(isinstance(p, torch.nn.Parameter) and p.requires_grad for p in m.parameters())
    m.optimizer = torch.optim.AdamW(m.parameters(), lr=1e-4)  # train_fusion_seq_level_decoder.py:410
    m.eval()
    assert not m.training
    m.train()
    assert m.cross_modal.layers[0].training
    assert m.double().emotion_decoder.emotion_queries.dtype == torch.float64


def test_errors_without_gpu():
    from hriemo.lib import HriemoError
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder

    m = FusionWithEmotionDecoder(d_model=64, n_heads=2, beta_hidden=16).eval()
    with pytest.raises(ValueError, match="Expected 2D or 3D tensor"):
        m(torch.zeros(1, 2, 3, 64), torch.zeros(2, 3, 64))
    with pytest.raises(HriemoError, match="no CPU fallback"):
        m(torch.zeros(2, 5, 64), torch.zeros(2, 3, 64))
    with pytest.raises(AssertionError):
        FusionWithEmotionDecoder(d_model=100, n_heads=8)


def test_fused_mask_rule():
    """_build_fused_mask (fusion_with_emotion_decoder.py:71-115): OR of the truncated audio mask and
    the text mask; a missing mask defers to the other; a shorter mask is PAD-extended."""
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
    import hriemo_oracle as O

    m = FusionWithEmotionDecoder(d_model=64, n_heads=2, beta_hidden=16)
    g = torch.Generator().manual_seed(3)
    ma, mt = O.ragged_masks(5, 20, g), O.ragged_masks(5, 8, g)
    for a, t, L in [(ma, mt, 8), (ma, None, 8), (None, mt, 8), (None, None, 8), (ma[:, :3], mt, 8)]:
        got, ref = m._build_fused_mask(a, t, L), O.build_fused_mask(a, t, L)
        assert (got is None and ref is None) or torch.equal(got, ref)
