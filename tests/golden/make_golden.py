#!/usr/bin/env python3
"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported read-only
from /root/reference) on seeded synthetic inputs.  Run in the authoring container:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as small
fixtures.  Weights are NOT stored for the full-size models (54.6 M parameters): a fixture
records the construction seed and a checksum of the state_dict; the drop-in modules
reproduce the reference's seeded initialisation exactly (tests/test_dropin_cpu.py), so the
tests rebuild identical weights from the seed and verify the checksum.  One tiny model is
stored with explicit weights and attention maps.
"""
import importlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
from hriemo_oracle import ragged_masks  # noqa: E402  (shared synthetic-mask generator)

REF = "/root/reference/models"
pkg = types.ModuleType("refmodels")
pkg.__path__ = [REF]
sys.modules["refmodels"] = pkg


def ref(mod):
    return importlib.import_module("refmodels." + mod)


def checksum(sd):
    return {"n_tensors": len(sd), "n_params": int(sum(v.numel() for v in sd.values())),
            "sum": float(sum(v.double().sum() for v in sd.values())),
            "abs_sum": float(sum(v.double().abs().sum() for v in sd.values()))}


def inputs(seed, B, T_a, T_t, d_a, d_t, masked):
    g = torch.Generator().manual_seed(seed)
    h_a = torch.randn(B, T_a, d_a, generator=g)
    h_t = torch.randn(B, T_t, d_t, generator=g)
    m_a = ragged_masks(B, T_a, g) if masked else None
    m_t = ragged_masks(B, T_t, g) if masked else None
    return h_a, h_t, m_a, m_t


def save(name, obj):
    path = os.path.join(HERE, name + ".pt")
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


@torch.no_grad()
def fusion_case(name, ctor_kwargs, model_seed, in_seed, B, T_a, T_t, masked, mosei=None):
    torch.manual_seed(model_seed)
    if mosei:
        m = ref("mosei_fusion_with_emotion_decoder").MoseiFusionWithEmotionDecoder(*mosei, **ctor_kwargs).eval()
        d_a, d_t = mosei
    else:
        m = ref("fusion_with_emotion_decoder").FusionWithEmotionDecoder(**ctor_kwargs).eval()
        d_a = d_t = ctor_kwargs.get("d_model", 768)
    h_a, h_t, m_a, m_t = inputs(in_seed, B, T_a, T_t, d_a, d_t, masked)
    logits, beta, z = m(h_a, h_t, m_a, m_t)
    save(name, dict(kind="mosei" if mosei else "fusion", ctor=ctor_kwargs, mosei=mosei, model_seed=model_seed,
                    in_seed=in_seed, B=B, T_a=T_a, T_t=T_t, masked=masked, weights=checksum(m.state_dict()),
                    logits=logits, beta=beta, z=z))


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    # BASELINE config 2: IEMOCAP seq-level defaults, T_a=300, T_t=50
    fusion_case("cfg2_iemocap_nomask", {}, 1234, 11, 4, 300, 50, False)
    fusion_case("cfg2_iemocap_ragged", {}, 1234, 12, 4, 300, 50, True)
    # north-star shape
    fusion_case("ns_500x64_ragged", {}, 1234, 13, 2, 500, 64, True)
    # utterance-level inputs through the seq-level model ([B,d] -> [B,1,d])
    torch.manual_seed(1234)
    m = ref("fusion_with_emotion_decoder").FusionWithEmotionDecoder().eval()
    g = torch.Generator().manual_seed(14)
    h_a, h_t = torch.randn(8, 768, generator=g), torch.randn(8, 768, generator=g)
    lo, be, z = m(h_a, h_t)
    save("utter_2d_inputs", dict(kind="fusion2d", model_seed=1234, in_seed=14, B=8, weights=checksum(m.state_dict()),
                                 logits=lo, beta=be, z=z))
    # BASELINE config 3: MOSEI wrapper (defaults and the v2 training variant)
    fusion_case("cfg3_mosei_default", {}, 1234, 21, 4, 300, 128, True, mosei=(74, 300))
    fusion_case("cfg3_mosei_v2", dict(num_layers_fusion=1, beta_hidden=64, dropout=0.4), 1234, 22, 4, 300, 128, True,
                mosei=(74, 300))
    # BASELINE config 1: utterance-level FusionClassifier, B=32 (and the seq-level shape of its test)
    torch.manual_seed(1234)
    fc = ref("fusion_classifier").FusionClassifier().eval()
    g = torch.Generator().manual_seed(31)
    h_a, h_t = torch.randn(32, 768, generator=g), torch.randn(32, 768, generator=g)
    lo, be, pooled = fc(h_a, h_t)
    h_a2, h_t2, m_a2, m_t2 = inputs(32, 4, 400, 128, 768, 768, True)
    lo2, be2, pooled2 = fc(h_a2, h_t2, m_a2, m_t2)
    save("cfg1_fusion_classifier", dict(kind="classifier", model_seed=1234, weights=checksum(fc.state_dict()),
                                        utter=dict(in_seed=31, B=32, logits=lo, beta=be, pooled=pooled),
                                        seq=dict(in_seed=32, B=4, T_a=400, T_t=128, logits=lo2, beta=be2, pooled=pooled2)))
    # reference tests/test_beta_gate.py and tests/test_cross_modal_block.py: legacy block + scalar gate
    torch.manual_seed(1234)
    cross = ref("cross_modal_block").CrossModalTransformer(num_layers=2, d_model=768, n_heads=8).eval()
    gate = ref("beta_gate").BetaGate(d_model=768, hidden_dim=256).eval()
    g = torch.Generator().manual_seed(41)
    h_a, h_t = torch.randn(32, 1, 768, generator=g), torch.randn(32, 1, 768, generator=g)
    a, t = cross(h_a, h_t)
    hf, beta = gate(a, t)
    g = torch.Generator().manual_seed(42)
    s_a, s_t = torch.randn(8, 400, 768, generator=g), torch.randn(8, 128, 768, generator=g)
    zm_a, zm_t = torch.zeros(8, 400, dtype=torch.bool), torch.zeros(8, 128, dtype=torch.bool)
    sa, st = cross(s_a, s_t, zm_a, zm_t)
    shf, sbeta = gate(sa, st, zm_a, zm_t)
    save("legacy_block_scalar_gate", dict(kind="legacy", model_seed=1234,
                                          weights_cross=checksum(cross.state_dict()), weights_gate=checksum(gate.state_dict()),
                                          utter=dict(in_seed=41, B=32, h_a_tilde=a, h_t_tilde=t, h_fusion=hf, beta=beta),
                                          seq=dict(in_seed=42, B=8, T_a=400, T_t=128,
                                                   h_a_tilde_slice=sa[:, ::40, ::32].clone(), h_t_tilde_slice=st[:, ::16, ::32].clone(),
                                                   h_fusion_slice=shf[:, ::16, ::32].clone(), beta=sbeta)))
    # tiny model with explicit weights, ragged masks and attention maps (pins the oracle incl. return_attention)
    torch.manual_seed(99)
    kw = dict(d_model=32, num_emotions=3, n_heads=1, num_layers_fusion=2, num_layers_decoder=2, beta_hidden=16, dropout=0.0)
    tiny = ref("fusion_with_emotion_decoder").FusionWithEmotionDecoder(**kw).eval()
    h_a, h_t, m_a, m_t = inputs(51, 3, 20, 9, 32, 32, True)
    lo, be, z, pack = tiny(h_a, h_t, m_a, m_t, return_attention=True)
    save("tiny_explicit_weights", dict(kind="tiny", ctor=kw, state_dict={k: v.clone() for k, v in tiny.state_dict().items()},
                                       h_a=h_a, h_t=h_t, mask_a=m_a, mask_t=m_t, logits=lo, beta=be, z=z, attn=pack))


if __name__ == "__main__":
    main()
