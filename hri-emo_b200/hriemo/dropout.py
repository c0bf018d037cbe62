"""Dropout of the training step: stream keys and the mask arithmetic of csrc/dropout.cuh restated in torch.

The reference applies nn.Dropout(p) to every sub-layer output and, through nn.MultiheadAttention(dropout=p), to the
attention probabilities (models/cross_modal_block_tacfn.py:24-38, 81-119; models/emotion_decoder.py:14-29, 42-59).
Here no mask is ever stored: element (row, col) of a stream is kept iff byte (col & 3) of
drop_word(key, row, col >> 2) is >= p8 = round(256 p); the backward pass recomputes the words from the same key.
A `Drop` object carries the rate and the seed of ONE forward / backward pair and hands every site its stream key.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

M32 = 0xFFFFFFFF
C_ROW, C_WORD, C_BH = 0x9E3779B1, 0x632BE5AB, 0xC2B2AE35


def mix(x: int) -> int:
    """lowbias32 on a Python int (csrc/dropout.cuh: drop_mix)."""
    x &= M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def _mix_t(x: torch.Tensor) -> torch.Tensor:
    x = x & M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & M32
    return x ^ (x >> 16)


def key_bh(key: int, bh: int) -> int:
    return mix(key + bh * C_BH)


def keep_mask(rows: int, cols: int, key: int, p8: int, device="cpu") -> torch.Tensor:
    """bool [rows, cols]: the keep mask of one stream (int64 torch arithmetic; the oracle of hriemo_dropout_mask)."""
    r = torch.arange(rows, dtype=torch.int64, device=device)[:, None]
    c = torch.arange(cols, dtype=torch.int64, device=device)[None, :]
    word = _mix_t(key + r * C_ROW + (c >> 2) * C_WORD)
    return ((word >> ((c & 3) * 8)) & 0xFF) >= p8


class Drop:
    """Rate + seed of one training forward / backward.  site ids: encoder layer i -> 1000 (i + 1) + kind, decoder
    layer i -> 100000 + 100 i + kind (kinds are listed where they are used, hriemo/backward.py)."""

    def __init__(self, p: float, seed: int):
        self.p8 = max(0, min(255, int(round(float(p) * 256.0))))
        self.scale = 1.0 / (1.0 - self.p8 / 256.0)
        self.seed = int(seed) & M32

    @property
    def on(self) -> bool:
        return self.p8 > 0

    def site(self, site_id: int) -> Optional[Tuple[int, float, int]]:
        """(p8, scale, key) of a site, or None when dropout is off (the wrappers then take their plain paths)."""
        if not self.on:
            return None
        return (self.p8, self.scale, mix(self.seed + site_id * C_ROW))


def make(p: float) -> Optional[Drop]:
    """A Drop for one step with a seed drawn from torch's CPU generator (reproducible under torch.manual_seed, no
    device synchronisation), or None for p = 0."""
    if p <= 0.0:
        return None
    d = Drop(p, int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
    return d if d.on else None
