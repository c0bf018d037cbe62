// Dropout of hidden-state tensors in the training step (nn.Dropout at models/cross_modal_block_tacfn.py:81, 92, 105, 106,
// 118, 119 and models/emotion_decoder.py:43, 55, 58, 59) and the mask itself for tests.  The mask is a function of
// (key, row, column) -- csrc/dropout.cuh -- so the backward pass calls the same kernel on the gradient with the same key.
#include <cuda_bf16.h>

#include "dropout.cuh"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

// out = (keep ? x * scale : 0) [+ resid]; one thread per four elements (one hash word)
template <bool F32>
__global__ void __launch_bounds__(256)
dropout_kernel(const void* __restrict__ x, int64_t ldx, const void* __restrict__ resid, int64_t ldr, void* __restrict__ out,
               int64_t ldo, int64_t rows, int cols, uint32_t p8, float scale, uint32_t key) {
  const int wpr = cols >> 2;   // words per row
  const int64_t total = rows * wpr;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / wpr;
    const int w = static_cast<int>(i - r * wpr);
    const uint32_t word = drop_word(key, static_cast<uint32_t>(r), static_cast<uint32_t>(w));
    float v[4], rs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (F32) {
      const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(x) + r * ldx + w * 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      if (resid != nullptr) {
        const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(resid) + r * ldr + w * 4);
        rs[0] = b.x; rs[1] = b.y; rs[2] = b.z; rs[3] = b.w;
      }
    } else {
      const uint2 a = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(x) + r * ldx + w * 4);
      v[0] = bf16_lo(a.x); v[1] = bf16_hi(a.x); v[2] = bf16_lo(a.y); v[3] = bf16_hi(a.y);
      if (resid != nullptr) {
        const uint2 b = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(resid) + r * ldr + w * 4);
        rs[0] = bf16_lo(b.x); rs[1] = bf16_hi(b.x); rs[2] = bf16_lo(b.y); rs[3] = bf16_hi(b.y);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (((word >> (e * 8)) & 0xffu) >= p8 ? v[e] * scale : 0.0f) + rs[e];
    if (F32) {
      *reinterpret_cast<float4*>(static_cast<float*>(out) + r * ldo + w * 4) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      uint2 o;
      o.x = pack_bf16(v[0], v[1]);
      o.y = pack_bf16(v[2], v[3]);
      *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + r * ldo + w * 4) = o;
    }
  }
}

// out[r, c] = 1 if element (row r of its stream, column c) is kept
__global__ void dropout_mask_kernel(uint8_t* __restrict__ out, int64_t rows, int cols, uint32_t key, uint32_t p8,
                                    int64_t rows_per_stream) {
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const uint32_t c = static_cast<uint32_t>(i - r * cols);
    uint32_t k = key, row = static_cast<uint32_t>(r);
    if (rows_per_stream > 0) {
      const int64_t s = r / rows_per_stream;
      k = drop_key_bh(key, static_cast<uint32_t>(s));
      row = static_cast<uint32_t>(r - s * rows_per_stream);
    }
    out[i] = drop_keep(k, row, c, p8) ? 1 : 0;
  }
}

}  // namespace hriemo

using namespace hriemo;

extern "C" int hriemo_dropout(const void* x, int32_t is_f32, int64_t ldx, const void* resid, int64_t ldr, void* out,
                              int64_t ldo, int64_t rows, int32_t cols, uint32_t p8, float scale, uint32_t key, void* stream) {
  HRIEMO_REQUIRE(x && out && rows >= 0 && cols > 0 && cols % 4 == 0, "dropout: bad argument (cols must be a multiple of 4)");
  HRIEMO_REQUIRE(p8 <= 255u, "dropout: p8=%u (round(256 p), at most 255)", p8);
  const int64_t al = is_f32 ? 4 : 4;   // elements per access: 16 bytes of f32, 8 bytes of bf16
  const uintptr_t mask = is_f32 ? 15u : 7u;
  HRIEMO_REQUIRE(ldx >= cols && ldo >= cols && ldx % al == 0 && ldo % al == 0 && (resid == nullptr || (ldr >= cols && ldr % al == 0)),
                 "dropout: bad leading dimension");
  HRIEMO_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(resid)) & mask) == 0,
                 "dropout: misaligned operand");
  if (rows == 0) return HRIEMO_OK;
  const int64_t total = rows * (cols / 4);
  int64_t grid = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  if (grid > cap) grid = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_f32)
    dropout_kernel<true><<<static_cast<unsigned>(grid), 256, 0, s>>>(x, ldx, resid, ldr, out, ldo, rows, cols, p8, scale, key);
  else
    dropout_kernel<false><<<static_cast<unsigned>(grid), 256, 0, s>>>(x, ldx, resid, ldr, out, ldo, rows, cols, p8, scale, key);
  return check_launch("dropout");
}

extern "C" int hriemo_dropout_mask(uint8_t* out, int64_t rows, int32_t cols, uint32_t key, uint32_t p8, int64_t rows_per_stream,
                                   void* stream) {
  HRIEMO_REQUIRE(out && rows >= 0 && cols > 0 && p8 <= 255u && rows_per_stream >= 0, "dropout_mask: bad argument");
  if (rows == 0) return HRIEMO_OK;
  const int64_t total = rows * cols;
  int64_t grid = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  if (grid > cap) grid = cap;
  dropout_mask_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, rows, cols, key, p8,
                                                                                                  rows_per_stream);
  return check_launch("dropout_mask");
}
