// Length-bucketed staging: the device side of the step BEFORE the path.
// The reference's collate zero-pads every utterance to the batch maximum and marks the tail with a
// True = PAD mask (scripts/fusion/train_fusion_seq_level_decoder.py:191-232,
// scripts/infer/mosei_eval_infer.py:128-147).  Padded rows never reach logits / beta / z (keys are
// masked, pooled means are masked, the decoder reads the fused mask), so the host sorts utterances by
// valid length and these kernels build one slab of similar-length utterances trimmed to the slab's
// own maximum: the GEMMs and attention then never see most of the padding.
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

static inline unsigned staging_grid(int64_t work_items, int threads) {
  const int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  return static_cast<unsigned>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

// lens[b] = index of the last valid (0) entry of pad[b, :] + 1, or 0 when every entry is PAD.  Warp per utterance.
__global__ void mask_lengths_kernel(const uint8_t* __restrict__ pad, int B, int T, int32_t* __restrict__ lens) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  int last = -1;
  for (int t = lane; t < T; t += 32)
    if (pad[static_cast<int64_t>(b) * T + t] == 0) last = t;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if (lane == 0) lens[b] = last + 1;
}

// out[i, t, c] = bf16(in[utt[i], t, c]) for t < min(T_in, T_out), c < cols; zero elsewhere (t >= T_in,
// cols <= c < ld_out).  One thread per 8 output elements (one 16-byte store), grid-stride.
template <bool F32>
__global__ void __launch_bounds__(256)
gather_utterances_kernel(const void* __restrict__ in, int64_t ld_in, int T_in, const int32_t* __restrict__ utt,
                         __nv_bfloat16* __restrict__ out, int64_t ld_out, int64_t n, int T_out, int cols,
                         int vec_ok) {
  const int64_t chunks_per_row = ld_out / 8;
  const int64_t total = n * T_out * chunks_per_row;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = idx / chunks_per_row;          // i * T_out + t
    const int c = static_cast<int>(idx - row * chunks_per_row) * 8;
    const int64_t i = row / T_out;
    const int t = static_cast<int>(row - i * T_out);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (t < T_in && c < cols) {
      const int64_t src_row = static_cast<int64_t>(__ldg(utt + i)) * T_in + t;
      if (F32) {
        const float* src = static_cast<const float*>(in) + src_row * ld_in + c;
        float f[8];
        if (vec_ok && c + 8 <= cols) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src));
          const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
          f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = (c + k < cols) ? __ldg(src + k) : 0.0f;
        }
        o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
        o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
      } else {
        const uint16_t* src = static_cast<const uint16_t*>(in) + src_row * ld_in + c;
        if (vec_ok && c + 8 <= cols) {
          o = __ldg(reinterpret_cast<const uint4*>(src));
        } else {
          uint16_t h[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) h[k] = (c + k < cols) ? __ldg(src + k) : static_cast<uint16_t>(0);
          o.x = h[0] | (static_cast<uint32_t>(h[1]) << 16); o.y = h[2] | (static_cast<uint32_t>(h[3]) << 16);
          o.z = h[4] | (static_cast<uint32_t>(h[5]) << 16); o.w = h[6] | (static_cast<uint32_t>(h[7]) << 16);
        }
      }
    }
    *reinterpret_cast<uint4*>(out + row * ld_out + c) = o;
  }
}

// out[i, t] = in[utt[i], t] for t < min(T_in, T_out), 1 (PAD) for T_in <= t < T_out.
__global__ void gather_masks_kernel(const uint8_t* __restrict__ in, int T_in, const int32_t* __restrict__ utt,
                                    uint8_t* __restrict__ out, int64_t n, int T_out) {
  const int64_t total = n * T_out;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = idx / T_out;
    const int t = static_cast<int>(idx - i * T_out);
    out[idx] = t < T_in ? in[static_cast<int64_t>(__ldg(utt + i)) * T_in + t] : static_cast<uint8_t>(1);
  }
}

// out[utt[i], :] = in[i, :] (f32 rows): puts a slab's results back at the utterances' original positions.
__global__ void scatter_rows_f32_kernel(const float* __restrict__ in, const int32_t* __restrict__ utt,
                                        float* __restrict__ out, int64_t n, int64_t cols) {
  const int64_t total = n * cols;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = idx / cols;
    out[static_cast<int64_t>(__ldg(utt + i)) * cols + (idx - i * cols)] = in[idx];
  }
}

}  // namespace hriemo

extern "C" int hriemo_mask_lengths(const uint8_t* pad, int32_t B, int32_t T, int32_t* lens, void* stream) {
  using namespace hriemo;
  HRIEMO_REQUIRE(pad && lens && B > 0 && T > 0, "mask_lengths: bad argument");
  mask_lengths_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(pad, B, T, lens);
  return check_launch("mask_lengths");
}

extern "C" int hriemo_gather_utterances_bf16(const void* in, int32_t in_is_f32, int64_t ld_in, int32_t T_in,
                                             const int32_t* utt, void* out_bf16, int64_t ld_out, int32_t n,
                                             int32_t T_out, int32_t cols, void* stream) {
  using namespace hriemo;
  HRIEMO_REQUIRE(in && utt && out_bf16, "gather_utterances: null pointer");
  HRIEMO_REQUIRE(n >= 0 && T_in > 0 && T_out > 0 && cols > 0 && ld_in >= cols && ld_out >= cols && ld_out % 8 == 0,
                 "gather_utterances: bad shape n=%d T_in=%d T_out=%d cols=%d ld_in=%lld ld_out=%lld", n, T_in, T_out,
                 cols, (long long)ld_in, (long long)ld_out);
  HRIEMO_REQUIRE((reinterpret_cast<uintptr_t>(out_bf16) & 15u) == 0, "gather_utterances: out misaligned");
  if (n == 0) return HRIEMO_OK;
  const int vec_ok = (reinterpret_cast<uintptr_t>(in) & 15u) == 0 && ld_in % (in_is_f32 ? 4 : 8) == 0;
  const int64_t total = static_cast<int64_t>(n) * T_out * (ld_out / 8);
  const unsigned grid = staging_grid(total, 256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(out_bf16);
  if (in_is_f32)
    gather_utterances_kernel<true><<<grid, 256, 0, s>>>(in, ld_in, T_in, utt, out, ld_out, n, T_out, cols, vec_ok);
  else
    gather_utterances_kernel<false><<<grid, 256, 0, s>>>(in, ld_in, T_in, utt, out, ld_out, n, T_out, cols, vec_ok);
  return check_launch("gather_utterances_bf16");
}

extern "C" int hriemo_gather_masks(const uint8_t* pad, int32_t T_in, const int32_t* utt, uint8_t* out, int32_t n,
                                   int32_t T_out, void* stream) {
  using namespace hriemo;
  HRIEMO_REQUIRE(pad && utt && out && n >= 0 && T_in > 0 && T_out > 0, "gather_masks: bad argument");
  if (n == 0) return HRIEMO_OK;
  const int64_t total = static_cast<int64_t>(n) * T_out;
  gather_masks_kernel<<<staging_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pad, T_in, utt, out, n,
                                                                                                T_out);
  return check_launch("gather_masks");
}

extern "C" int hriemo_scatter_rows_f32(const float* in, const int32_t* utt, float* out, int64_t n, int64_t cols,
                                       void* stream) {
  using namespace hriemo;
  HRIEMO_REQUIRE(in && utt && out && n >= 0 && cols > 0, "scatter_rows: bad argument");
  if (n == 0) return HRIEMO_OK;
  scatter_rows_f32_kernel<<<staging_grid(n * cols, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, utt, out, n,
                                                                                                       cols);
  return check_launch("scatter_rows_f32");
}
