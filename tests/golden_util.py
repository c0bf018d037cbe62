"""Shared helpers: load a golden fixture, rebuild the (seeded) model and inputs it describes."""
import os

import torch

from hriemo_oracle import ragged_masks

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def checksum(sd):
    return {"n_tensors": len(sd), "n_params": int(sum(v.numel() for v in sd.values())),
            "sum": float(sum(v.double().sum() for v in sd.values())),
            "abs_sum": float(sum(v.double().abs().sum() for v in sd.values()))}


def assert_same_weights(module, expected):
    got = checksum(module.state_dict())
    assert got["n_tensors"] == expected["n_tensors"] and got["n_params"] == expected["n_params"], (got, expected)
    assert abs(got["sum"] - expected["sum"]) <= 1e-9 * max(1.0, abs(expected["abs_sum"])), (got, expected)
    assert abs(got["abs_sum"] - expected["abs_sum"]) <= 1e-9 * expected["abs_sum"], (got, expected)


def make_inputs(seed, B, T_a, T_t, d_a, d_t, masked):
    g = torch.Generator().manual_seed(seed)
    h_a = torch.randn(B, T_a, d_a, generator=g)
    h_t = torch.randn(B, T_t, d_t, generator=g)
    m_a = ragged_masks(B, T_a, g) if masked else None
    m_t = ragged_masks(B, T_t, g) if masked else None
    return h_a, h_t, m_a, m_t


def build_fusion(fx):
    """Drop-in model with the fixture's seeded random-init weights (checksum-verified) + inputs."""
    from models.fusion_with_emotion_decoder import FusionWithEmotionDecoder
    from models.mosei_fusion_with_emotion_decoder import MoseiFusionWithEmotionDecoder

    torch.manual_seed(fx["model_seed"])
    if fx["kind"] == "mosei":
        m = MoseiFusionWithEmotionDecoder(*fx["mosei"], **fx["ctor"]).eval()
        d_a, d_t = fx["mosei"]
    else:
        m = FusionWithEmotionDecoder(**fx.get("ctor", {})).eval()
        d_a = d_t = fx.get("ctor", {}).get("d_model", 768)
    assert_same_weights(m, fx["weights"])
    if fx["kind"] == "fusion2d":
        g = torch.Generator().manual_seed(fx["in_seed"])
        ins = (torch.randn(fx["B"], 768, generator=g), torch.randn(fx["B"], 768, generator=g), None, None)
    else:
        ins = make_inputs(fx["in_seed"], fx["B"], fx["T_a"], fx["T_t"], d_a, d_t, fx["masked"])
    return m, ins


def n_heads_of(fx):
    if fx["kind"] == "mosei":
        return fx["ctor"].get("n_heads", 4)
    return fx.get("ctor", {}).get("n_heads", 8)


def to_dev(x, dev):
    return None if x is None else x.to(dev)
