"""The stock-PyTorch competitor arm (baseline/stock_torch.py) is the reference's module graph: on the CPU it must
reproduce the golden outputs of the unmodified reference (tests/golden/*.pt) to float32 round-off."""
import pytest
import torch

import golden_util as G
from baseline.stock_torch import StockFusion, run_mode


@pytest.mark.parametrize("name", ["cfg2_iemocap_ragged", "cfg2_iemocap_nomask", "cfg3_mosei_default", "cfg3_mosei_v2"])
def test_stock_torch_matches_reference_outputs(name):
    fx = G.load(name)
    model, (h_a, h_t, m_a, m_t) = G.build_fusion(fx)
    stock = StockFusion(model.state_dict(), G.n_heads_of(fx)).eval()
    lo, be, z = run_mode(stock, "fp32", h_a, h_t, m_a, m_t)
    assert lo.shape == fx["logits"].shape and be.shape == fx["beta"].shape and z.shape == fx["z"].shape
    assert (lo - fx["logits"]).abs().max().item() < 2e-6
    assert (be - fx["beta"]).abs().max().item() < 2e-6
    assert (z - fx["z"]).abs().max().item() < 1e-5
