// Counter-based dropout masks shared by every kernel that applies or re-applies dropout in the training step
// (nn.Dropout / nn.MultiheadAttention(dropout=p) of the reference: models/cross_modal_block_tacfn.py:24-38, 81-119,
// models/emotion_decoder.py:14-29, 42-59).  No mask is ever stored: element (row, col) of a stream is kept iff byte
// (col & 3) of  drop_word(key, row, col >> 2)  is >= p8, where p8 = round(256 p) and key identifies the stream (the site
// in the network and the optimizer step; per (utterance, head) for attention probabilities, see drop_key_bh).  The
// backward pass recomputes the same words from the same key -- in either orientation: a thread that owns a row walks
// the words of its row (one IMAD + one mix per four elements), a thread that owns a column walks the rows.
// Kept elements are scaled by 1 / (1 - p8 / 256) (the caller passes the scale), so the expectation is exact for the
// quantised rate.  hriemo/dropout.py holds the same arithmetic in torch for the tests.
#pragma once
#include <stdint.h>

namespace hriemo {

constexpr uint32_t DROP_C_ROW = 0x9E3779B1u;
constexpr uint32_t DROP_C_WORD = 0x632BE5ABu;
constexpr uint32_t DROP_C_BH = 0xC2B2AE35u;

// lowbias32 (full-avalanche 32-bit mixer)
__host__ __device__ __forceinline__ uint32_t drop_mix(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// four keep-bytes for elements (row, 4 w .. 4 w + 3)
__host__ __device__ __forceinline__ uint32_t drop_word(uint32_t key, uint32_t row, uint32_t w) {
  return drop_mix(key + row * DROP_C_ROW + w * DROP_C_WORD);
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t key, uint32_t row, uint32_t col, uint32_t p8) {
  return ((drop_word(key, row, col >> 2) >> ((col & 3u) * 8u)) & 0xffu) >= p8;
}
// the stream of attention probabilities of one (utterance, head): rows are queries, columns keys
__host__ __device__ __forceinline__ uint32_t drop_key_bh(uint32_t key, uint32_t bh) { return drop_mix(key + bh * DROP_C_BH); }

// 32-bit lane masks for two packed bf16 values (elements 2 i, 2 i + 1 of a group of four): 0xffff per kept half
__device__ __forceinline__ uint32_t drop_pair_mask(uint32_t word, int pair, uint32_t p8) {
  const uint32_t b0 = (word >> (pair * 16)) & 0xffu, b1 = (word >> (pair * 16 + 8)) & 0xffu;
  return (b0 >= p8 ? 0x0000ffffu : 0u) | (b1 >= p8 ? 0xffff0000u : 0u);
}

}  // namespace hriemo
