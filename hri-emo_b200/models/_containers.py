"""Parameter containers with the reference's names, shapes and initialisation.

The reference stores its attention weights in ``nn.MultiheadAttention`` modules;
only their *parameters* matter for the drop-in contract (state_dict keys
``in_proj_weight``, ``in_proj_bias``, ``out_proj.weight``, ``out_proj.bias`` — SURVEY
Appendix B).  ``MHAParams`` reproduces that parameter set and PyTorch's
initialisation order (out_proj Linear init, then xavier-uniform in-projection, zero
biases) so that ``torch.manual_seed(s)`` gives the same random-init weights as the
reference.  ``nn.Linear`` / ``nn.LayerNorm`` / ``nn.Sequential`` are used below purely
as parameter holders; their ``forward`` is never called.
"""
import torch
import torch.nn as nn


class MHAParams(nn.Module):
    def __init__(self, embed_dim: int, num_heads: int, dropout: float = 0.0):
        super().__init__()
        if embed_dim % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        self.dropout = dropout
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.in_proj_bias, 0.0)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, *args, **kwargs):  # pragma: no cover
        raise RuntimeError("MHAParams is a parameter container; attention runs in the fused kernels")


def ffn(d_model: int, d_hidden: int) -> nn.Sequential:
    """Linear -> ReLU -> Linear holder (keys '0.*' and '2.*')."""
    return nn.Sequential(nn.Linear(d_model, d_hidden), nn.ReLU(), nn.Linear(d_hidden, d_model))
