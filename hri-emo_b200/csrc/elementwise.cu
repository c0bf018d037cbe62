// HBM-bound passes of the path: feature cast, LayerNorm, the beta-gate pooling
// and blend.  All are one-warp-per-row kernels with 16-byte vector accesses and
// fp32 statistics; none needs tensor cores.
#include <stdlib.h>

#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

constexpr int ROW_MAX_VEC = 8;  // 8 lanes-iterations x 32 lanes x 8 elements -> d <= 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
  u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}

// A row of d elements held by one warp: lane l owns elements [(i*32 + l)*8, +8) for i < NV,
// NV = ceil(d / 256) chosen at launch (template) so the register footprint matches d.
template <int NV>
struct RowRegs {
  float v[NV][8];
};

template <bool F32, int NV>
__device__ __forceinline__ void load_row(RowRegs<NV>& r, const void* row, int d, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
      if (F32) {
        const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(row) + c);
        const float4 a = __ldg(p), b = __ldg(p + 1);
        r.v[i][0] = a.x; r.v[i][1] = a.y; r.v[i][2] = a.z; r.v[i][3] = a.w;
        r.v[i][4] = b.x; r.v[i][5] = b.y; r.v[i][6] = b.z; r.v[i][7] = b.w;
      } else {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(row) + c));
        unpack8(u, r.v[i]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) r.v[i][k] = 0.0f;
    }
  }
}

// In-place LayerNorm of a warp-held row (two-pass: mean, then biased variance).
template <int NV>
__device__ __forceinline__ void normalize_row(RowRegs<NV>& r, int d, int lane, const float* gamma,
                                              const float* beta, float eps) {
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) s += r.v[i][k];
  const float mean = warp_sum(s) / static_cast<float>(d);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t = r.v[i][k] - mean;
        q += t * t;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(d) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) r.v[i][k] = (r.v[i][k] - mean) * rstd * g[k] + bb[k];
    }
  }
}

// (x - mean) * rstd * gamma + beta with the row statistics already known (written by the GEMM
// epilogue that produced x): no reduction.
template <int NV>
__device__ __forceinline__ void affine_row(RowRegs<NV>& r, int d, int lane, const float* gamma, const float* beta,
                                           float mean, float rstd) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) r.v[i][k] = (r.v[i][k] - mean) * rstd * g[k] + bb[k];
    }
  }
}

template <int NV>
__device__ __forceinline__ void store_row(const RowRegs<NV>& r, __nv_bfloat16* yb, float* yf, int d, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
      if (yb) *reinterpret_cast<uint4*>(yb + c) = pack8(r.v[i]);
      if (yf) {
        float4* p = reinterpret_cast<float4*>(yf + c);
        p[0] = make_float4(r.v[i][0], r.v[i][1], r.v[i][2], r.v[i][3]);
        p[1] = make_float4(r.v[i][4], r.v[i][5], r.v[i][6], r.v[i][7]);
      }
    }
  }
}

// ------------------------------------------------------------------ cast
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, int64_t ld_in,
                                     __nv_bfloat16* __restrict__ out, int64_t ld_out, int64_t rows,
                                     int cols, int vec_ok) {
  const int64_t chunks_per_row = ld_out / 8;
  const int64_t total = rows * chunks_per_row;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = idx / chunks_per_row;
    const int c = static_cast<int>(idx - row * chunks_per_row) * 8;
    const float* src = in + row * ld_in + c;
    float f[8];
    if (vec_ok && c + 8 <= cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = (c + k < cols) ? __ldg(src + k) : 0.0f;
    }
    *reinterpret_cast<uint4*>(out + row * ld_out + c) = pack8(f);
  }
}

// ------------------------------------------------------------------ LayerNorm
template <bool F32, int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ yb,
                 float* __restrict__ yf, int64_t ldy, int64_t rows, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  RowRegs<NV> r;
  const void* xr = F32 ? static_cast<const void*>(static_cast<const float*>(x) + row * ldx)
                       : static_cast<const void*>(static_cast<const __nv_bfloat16*>(x) + row * ldx);
  load_row<F32, NV>(r, xr, d, lane);
  normalize_row(r, d, lane, gamma, beta, eps);
  store_row(r, yb ? yb + row * ldy : nullptr, yf ? yf + row * ldy : nullptr, d, lane);
}

// ------------------------------------------------------------------ gate: LN + masked mean
// Row statistics of a warp-held row (two-pass: mean, then biased variance).
template <int NV>
__device__ __forceinline__ void row_stats(const RowRegs<NV>& r, int d, int lane, float eps, float& mean, float& rstd) {
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) s += r.v[i][k];
  mean = warp_sum(s) / static_cast<float>(d);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t = r.v[i][k] - mean;
        q += t * t;
      }
    }
  }
  rstd = rsqrtf(warp_sum(q) / static_cast<float>(d) + eps);
}

// One CTA per utterance; warp w reduces rows t = w, w+8, ... and the 8 partial sums are combined through shared memory in
// a fixed order (deterministic).
//  * Rows reach the warp through a warp-PRIVATE ring of LMM stages in shared memory filled with 16-byte cp.async copies:
//    every lane copies exactly the bytes it later reads, so the ring needs no barrier, and STAGES - 1 rows (not one) are in
//    flight per warp -- the register-prefetch form kept 24 KB in flight per SM and ran at 0.36 of the HBM roofline.
//  * The arithmetic is packed fp32 (FFMA2 / FADD2) with one-pass statistics (sum and sum of squares reduced together: one
//    round of shuffles per row instead of two).  The pending LayerNorm's (gamma, beta) live in registers for the whole
//    utterance; the gate's own LayerNorm is linear after the row statistics, so its gamma / beta are applied once to the
//    sum:  mean_t LN(y_t) = gamma * [sum_t (y_t - m_t) rstd_t] / n + beta * count / n,  and the "- m_t rstd_t" term is a
//    per-row SCALAR accumulated on the side.
constexpr int LMM_WARPS = 8;
template <int NV> struct LmmStages { static constexpr int value = NV <= 3 ? 6 : (NV == 4 ? 5 : 3); };

__device__ __forceinline__ void lmm_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int NV>
__global__ void __launch_bounds__(256)
ln_masked_mean_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, int apply_ln,
                      const uint8_t* __restrict__ pad, float* __restrict__ pooled, int64_t ld_pooled,
                      int T, int d, const float* __restrict__ pre_g, const float* __restrict__ pre_b,
                      const float2* __restrict__ pre_stats) {
  constexpr int STAGES = LmmStages<NV>::value;
  constexpr int ROW_BYTES = NV * 512;
  extern __shared__ __align__(16) uint8_t lmm_smem[];   // ring [8 warps][STAGES][NV * 32 lanes][16 B]; later part[8][d]
  float* part = reinterpret_cast<float*>(lmm_smem);
  // the pending LayerNorm's (mean, rstd) of a row travel with the row: an 8-byte cp.async per stage.  (Fetched with a
  // plain load at the point of use they were a dependent ~1 000-cycle DRAM round trip in every row's chain: 0.50 of HBM.)
  __shared__ __align__(8) float2 lmm_stats[LMM_WARPS][8];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = smem_u32(lmm_smem) + warp * (STAGES * ROW_BYTES) + lane * 16;
  const uint8_t* ring_p = lmm_smem + warp * (STAGES * ROW_BYTES) + lane * 16;
  float2 acc[NV][4], pg[NV][4], pb[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[i][k] = make_float2(0.0f, 0.0f);
      const bool on = pre_g != nullptr && c < d;
      pg[i][k] = on ? make_float2(__ldg(pre_g + c + 2 * k), __ldg(pre_g + c + 2 * k + 1)) : make_float2(0.0f, 0.0f);
      pb[i][k] = on ? make_float2(__ldg(pre_b + c + 2 * k), __ldg(pre_b + c + 2 * k + 1)) : make_float2(0.0f, 0.0f);
    }
  }
  int count = 0;
  float side = 0.0f;   // sum over rows of m_t * rstd_t (subtracted from every column at the end)
  // the utterance's PAD bytes in shared memory (read from global memory they were one dependent L2 round trip per row in
  // the issue chain of a masked batch); longer sequences keep reading global memory
  __shared__ uint8_t lmm_pad[2048];
  const bool pad_smem = pad != nullptr && T <= 2048;
  if (pad_smem) {
    for (int i = threadIdx.x; i < T; i += blockDim.x) lmm_pad[i] = pad[static_cast<int64_t>(b) * T + i];
    __syncthreads();
  }
  auto next_valid = [&](int t) {
    if (pad_smem) {
      while (t < T && lmm_pad[t] != 0) t += LMM_WARPS;   // warp-uniform
    } else {
      while (t < T && pad != nullptr && pad[static_cast<int64_t>(b) * T + t] != 0) t += LMM_WARPS;
    }
    return t;
  };
  auto issue = [&](int t, int slot) {   // one commit group per call, also when there is nothing left to copy
    if (t < T) {
      const __nv_bfloat16* row = x + (static_cast<int64_t>(b) * T + t) * ldx;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        if (c < d) lmm_cp_async16(ring + slot * ROW_BYTES + i * 512, row + c);
      }
      if (pre_stats != nullptr && lane == 0)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&lmm_stats[warp][slot])),
                     "l"(pre_stats + static_cast<int64_t>(b) * T + t) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int t = next_valid(warp);
  int t_iss = t;
#pragma unroll
  for (int s0 = 0; s0 < STAGES - 1; ++s0) {
    issue(t_iss, s0);
    if (t_iss < T) t_iss = next_valid(t_iss + LMM_WARPS);
  }
  const float inv_d = 1.0f / static_cast<float>(d);
  int slot = 0;
  while (t < T) {
    ++count;
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
    float2 st_row = make_float2(0.0f, 1.0f);
    if (pre_stats != nullptr) {
      __syncwarp();   // lane 0's copy of the statistics is complete (its wait_group above) and visible to the warp
      st_row = lmm_stats[warp][slot];
    }
    float2 r[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (c < d) u = *reinterpret_cast<const uint4*>(ring_p + slot * ROW_BYTES + i * 512);
      r[i][0] = make_float2(bf16_lo(u.x), bf16_hi(u.x));
      r[i][1] = make_float2(bf16_lo(u.y), bf16_hi(u.y));
      r[i][2] = make_float2(bf16_lo(u.z), bf16_hi(u.z));
      r[i][3] = make_float2(bf16_lo(u.w), bf16_hi(u.w));
    }
    // the slot consumed in the PREVIOUS iteration takes the row STAGES - 1 ahead
    issue(t_iss, slot == 0 ? STAGES - 1 : slot - 1);
    if (t_iss < T) t_iss = next_valid(t_iss + LMM_WARPS);
    if (pre_g != nullptr) {
      float m1, r1;
      if (pre_stats != nullptr) {
        m1 = st_row.x;
        r1 = st_row.y;
      } else {   // two-pass statistics of the raw row (only streams whose producer kept no statistics)
        float s1 = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) s1 += r[i][k].x + r[i][k].y;
        m1 = warp_sum(s1) * inv_d;
        float q1 = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if ((i * 32 + lane) * 8 < d) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float a0 = r[i][k].x - m1, a1 = r[i][k].y - m1;
              q1 += a0 * a0 + a1 * a1;
            }
          }
        }
        r1 = rsqrtf(warp_sum(q1) * inv_d + eps);
      }
      const float2 u2 = make_float2(r1, r1), k2 = make_float2(-m1 * r1, -m1 * r1);
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) r[i][k] = ffma2(ffma2(r[i][k], u2, k2), pg[i][k], pb[i][k]);   // 0 past d (pg = pb = 0)
    }
    if (apply_ln) {
      float2 sy = make_float2(0.0f, 0.0f), sq = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sy = fadd2(sy, r[i][k]);
          sq = ffma2(r[i][k], r[i][k], sq);
        }
      float s1 = sy.x + sy.y, s2 = sq.x + sq.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const float m2 = s1 * inv_d;
      const float r2 = rsqrtf(fmaxf(s2 * inv_d - m2 * m2, 0.0f) + eps);
      side = fmaf(m2, r2, side);
      const float2 r22 = make_float2(r2, r2);
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = ffma2(r[i][k], r22, acc[i][k]);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = fadd2(acc[i][k], r[i][k]);
    }
    slot = slot + 1 == STAGES ? 0 : slot + 1;
    t = next_valid(t + LMM_WARPS);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __shared__ int counts[LMM_WARPS];
  if (lane == 0) counts[warp] = count;
  __syncthreads();   // every warp is done with its ring: the memory becomes part[8][d]
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        part[warp * d + c + 2 * k] = acc[i][k].x - side;
        part[warp * d + c + 2 * k + 1] = acc[i][k].y - side;
      }
    }
  }
  __syncthreads();
  int total = 0;
#pragma unroll
  for (int w = 0; w < LMM_WARPS; ++w) total += counts[w];
  // mask None -> plain mean over T; else sum / clamp(count, 1)   (beta_gate_tacfn.py:17-24)
  const float denom = pad == nullptr ? static_cast<float>(T) : fmaxf(static_cast<float>(total), 1.0f);
  const float frac = static_cast<float>(total) / denom;   // 1, or 0 when every row is PAD
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < LMM_WARPS; ++w) s += part[w * d + c];
    s /= denom;
    pooled[static_cast<int64_t>(b) * ld_pooled + c] = apply_ln ? fmaf(__ldg(gamma + c), s, __ldg(beta + c) * frac) : s;
  }
}

__global__ void gate_input_kernel(const float* __restrict__ a, const float* __restrict__ t,
                                  float* __restrict__ g, int B, int d) {
  const int64_t total = static_cast<int64_t>(B) * d;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = idx / d;
    const int c = static_cast<int>(idx - b * d);
    const float av = a[idx], tv = t[idx];
    float* gr = g + b * 4 * d;
    gr[c] = av;
    gr[d + c] = tv;
    gr[2 * d + c] = fabsf(av - tv);
    gr[3 * d + c] = av * tv;
  }
}

// ------------------------------------------------------------------ gate: blend
template <int NV>
__global__ void __launch_bounds__(256)
gate_blend_kernel(const __nv_bfloat16* __restrict__ a, int64_t lda, int T_a,
                  const __nv_bfloat16* __restrict__ t, int64_t ldt, const float* __restrict__ ga,
                  const float* __restrict__ ba, const float* __restrict__ gt, const float* __restrict__ bt,
                  float eps, int apply_ln, const float* __restrict__ w, int w_is_scalar,
                  __nv_bfloat16* __restrict__ hb, float* __restrict__ hf, int64_t ldh,
                  float* __restrict__ beta_out, int B, int L, int d, const float* __restrict__ pga,
                  const float* __restrict__ pba, const float* __restrict__ pgt, const float* __restrict__ pbt,
                  const float2* __restrict__ psa, const float2* __restrict__ pst) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= static_cast<int64_t>(B) * L) return;
  const int b = static_cast<int>(row / L);
  const int tt = static_cast<int>(row - static_cast<int64_t>(b) * L);
  RowRegs<NV> ra, rt;
  load_row<false, NV>(ra, a + (static_cast<int64_t>(b) * T_a + tt) * lda, d, lane);
  load_row<false, NV>(rt, t + (static_cast<int64_t>(b) * L + tt) * ldt, d, lane);
  if (pga != nullptr) {
    if (psa != nullptr) {
      const float2 st = __ldg(psa + static_cast<int64_t>(b) * T_a + tt);
      affine_row(ra, d, lane, pga, pba, st.x, st.y);
    } else {
      normalize_row(ra, d, lane, pga, pba, eps);
    }
  }
  if (pgt != nullptr) {
    if (pst != nullptr) {
      const float2 st = __ldg(pst + static_cast<int64_t>(b) * L + tt);
      affine_row(rt, d, lane, pgt, pbt, st.x, st.y);
    } else {
      normalize_row(rt, d, lane, pgt, pbt, eps);
    }
  }
  if (apply_ln) {
    normalize_row(ra, d, lane, ga, ba, eps);
    normalize_row(rt, d, lane, gt, bt, eps);
  }
  float wsum = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float wv = w_is_scalar ? __ldg(w + b) : __ldg(w + static_cast<int64_t>(b) * d + c + k);
        wsum += wv;
        ra.v[i][k] = wv * ra.v[i][k] + (1.0f - wv) * rt.v[i][k];
      }
    }
  }
  store_row(ra, hb ? hb + row * ldh : nullptr, hf ? hf + row * ldh : nullptr, d, lane);
  if (tt == 0 && beta_out != nullptr) {
    wsum = warp_sum(wsum);
    if (lane == 0) beta_out[b] = w_is_scalar ? __ldg(w + b) : wsum / static_cast<float>(d);
  }
}

// The blend, streaming form (default whenever pending LayerNorms come with their row statistics, as in the forward).
// gate_blend_kernel above gives every row its own short-lived warp: two 1.5 KB loads in flight per warp, then ~1 500
// instructions (four two-pass LayerNorms with their gamma / beta re-read from L1 for every row, scalar gate loads):
// 566 MB in 340 us = 0.25 of the HBM roofline (ncu launch list v17).  Here one persistent CTA per SM; a warp owns a
// contiguous range of rows and a private cp.async ring (STAGES row pairs of a | t in flight per warp), and per row does
//     x = pending LayerNorm (known statistics; gamma / beta staged once per CTA in lane order in shared memory)
//     (s, q) = one-pass sums of both streams, reduced together (one round of shuffles for four sums)
//     h = A o (x_a r_a + k_a) + T o (x_t r_t + k_t) + C,   A = w o gamma_a,  T = (1 - w) o gamma_t,  C = w o beta_a + (1 - w) o beta_t
// with (A, T, C) folded once per utterance into registers -- all in packed fp32 (FFMA2).  ~700 instructions per row.
template <int NV> struct GbStages { static constexpr int value = NV <= 3 ? 6 : (NV == 4 ? 4 : 2); };

template <int NV>
__global__ void __launch_bounds__(256, 1)
gate_blend_stream_kernel(const __nv_bfloat16* __restrict__ a, int64_t lda, int T_a,
                         const __nv_bfloat16* __restrict__ t, int64_t ldt, const float* __restrict__ ga,
                         const float* __restrict__ ba, const float* __restrict__ gt, const float* __restrict__ bt,
                         float eps, int apply_ln, const float* __restrict__ w, int w_is_scalar,
                         __nv_bfloat16* __restrict__ hb, float* __restrict__ hf, int64_t ldh,
                         float* __restrict__ beta_out, int B, int L, int d, const float* __restrict__ pga,
                         const float* __restrict__ pba, const float* __restrict__ pgt, const float* __restrict__ pbt,
                         const float2* __restrict__ psa, const float2* __restrict__ pst) {
  constexpr int STAGES = GbStages<NV>::value;
  constexpr int PAIR_BYTES = 2 * NV * 512;
  extern __shared__ __align__(16) uint8_t gb_smem[];
  float4* vec = reinterpret_cast<float4*>(gb_smem);   // [4 vectors][NV][2 halves][32 lanes]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const float* srcs[4] = {pga, pba, pgt, pbt};
    for (int e = threadIdx.x; e < 4 * NV * 64; e += blockDim.x) {
      const int vsel = e / (NV * 64), rem = e - vsel * (NV * 64);
      const int i = rem >> 6, half = (rem >> 5) & 1, ln = rem & 31;
      const int c = (i * 32 + ln) * 8 + half * 4;
      float4 val = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (srcs[vsel] != nullptr && c < d) val = __ldg(reinterpret_cast<const float4*>(srcs[vsel] + c));
      vec[e] = val;
    }
  }
  __syncthreads();
  uint8_t* ring_p = gb_smem + 4 * NV * 64 * 16 + warp * (STAGES * PAIR_BYTES) + lane * 16;
  const uint32_t ring = smem_u32(ring_p);
  const int64_t rows = static_cast<int64_t>(B) * L;
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * 8 + warp, GW = static_cast<int64_t>(gridDim.x) * 8;
  const int64_t r_begin = rows * gw / GW, r_end = rows * (gw + 1) / GW;
  auto issue = [&](int64_t r, int slot) {   // one commit group per call, also when there is nothing left to copy
    if (r < r_end) {
      const int64_t b = r / L;
      const int tt = static_cast<int>(r - b * L);
      const __nv_bfloat16* ar = a + (b * T_a + tt) * lda;
      const __nv_bfloat16* tr = t + r * ldt;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        if (c < d) {
          lmm_cp_async16(ring + slot * PAIR_BYTES + i * 512, ar + c);
          lmm_cp_async16(ring + slot * PAIR_BYTES + (NV + i) * 512, tr + c);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto row_stats_of = [&](int64_t r, float2& sa, float2& st) {   // (mean, rstd) of the pending LayerNorms of row r
    sa = make_float2(0.0f, 1.0f);
    st = make_float2(0.0f, 1.0f);
    if (r < r_end) {
      const int64_t b = r / L;
      if (psa != nullptr) sa = __ldg(psa + b * T_a + (r - b * L));
      if (pst != nullptr) st = __ldg(pst + r);
    }
  };
#pragma unroll
  for (int s0 = 0; s0 < STAGES - 1; ++s0) issue(r_begin + s0, s0);
  float2 nsa, nst;
  row_stats_of(r_begin, nsa, nst);
  float2 cA[NV][4], cT[NV][4], cC[NV][4];
  int64_t b_cur = -1;
  const float inv_d = 1.0f / static_cast<float>(d);
  int slot = 0;
  for (int64_t r = r_begin; r < r_end; ++r) {
    const int64_t b = r / L;
    const float2 sta = nsa, stt = nst;
    row_stats_of(r + 1, nsa, nst);   // consumed one row later
    if (b != b_cur) {
      b_cur = b;
      float wsum = 0.0f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        float wv[8], g1[8], b1[8], g2[8], b2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { wv[k] = 0.0f; g1[k] = g2[k] = 1.0f; b1[k] = b2[k] = 0.0f; }
        if (c < d) {
          if (w_is_scalar) {
            const float ws = __ldg(w + b);
#pragma unroll
            for (int k = 0; k < 8; ++k) wv[k] = ws;
          } else {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + b * d + c)), w1 = __ldg(reinterpret_cast<const float4*>(w + b * d + c + 4));
            wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w; wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
          }
          if (apply_ln) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { g1[k] = __ldg(ga + c + k); b1[k] = __ldg(ba + c + k); g2[k] = __ldg(gt + c + k); b2[k] = __ldg(bt + c + k); }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float w0 = wv[2 * k], w1 = wv[2 * k + 1];
          wsum += w0 + w1;
          cA[i][k] = make_float2(w0 * g1[2 * k], w1 * g1[2 * k + 1]);
          cT[i][k] = make_float2((1.0f - w0) * g2[2 * k], (1.0f - w1) * g2[2 * k + 1]);
          cC[i][k] = make_float2(w0 * b1[2 * k] + (1.0f - w0) * b2[2 * k], w1 * b1[2 * k + 1] + (1.0f - w1) * b2[2 * k + 1]);
          if (c >= d) cA[i][k] = cT[i][k] = cC[i][k] = make_float2(0.0f, 0.0f);
        }
      }
      if (beta_out != nullptr && r == b * L) {   // this warp holds the utterance's first row
        wsum = warp_sum(wsum);
        if (lane == 0) beta_out[b] = w_is_scalar ? __ldg(w + b) : wsum * inv_d;
      }
    }
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
    float2 xa[NV][4], xt[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      uint4 u = make_uint4(0u, 0u, 0u, 0u), v = make_uint4(0u, 0u, 0u, 0u);
      if (c < d) {
        u = *reinterpret_cast<const uint4*>(ring_p + slot * PAIR_BYTES + i * 512);
        v = *reinterpret_cast<const uint4*>(ring_p + slot * PAIR_BYTES + (NV + i) * 512);
      }
      xa[i][0] = make_float2(bf16_lo(u.x), bf16_hi(u.x)); xa[i][1] = make_float2(bf16_lo(u.y), bf16_hi(u.y));
      xa[i][2] = make_float2(bf16_lo(u.z), bf16_hi(u.z)); xa[i][3] = make_float2(bf16_lo(u.w), bf16_hi(u.w));
      xt[i][0] = make_float2(bf16_lo(v.x), bf16_hi(v.x)); xt[i][1] = make_float2(bf16_lo(v.y), bf16_hi(v.y));
      xt[i][2] = make_float2(bf16_lo(v.z), bf16_hi(v.z)); xt[i][3] = make_float2(bf16_lo(v.w), bf16_hi(v.w));
    }
    // the slot consumed in the PREVIOUS iteration takes the row STAGES - 1 ahead
    issue(r + STAGES - 1, slot == 0 ? STAGES - 1 : slot - 1);
    if (pga != nullptr) {
      const float2 u2 = make_float2(sta.y, sta.y), k2 = make_float2(-sta.x * sta.y, -sta.x * sta.y);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 g0 = vec[(0 * NV + i) * 64 + lane], g1 = vec[(0 * NV + i) * 64 + 32 + lane];
        const float4 b0 = vec[(1 * NV + i) * 64 + lane], b1 = vec[(1 * NV + i) * 64 + 32 + lane];
        xa[i][0] = ffma2(ffma2(xa[i][0], u2, k2), make_float2(g0.x, g0.y), make_float2(b0.x, b0.y));
        xa[i][1] = ffma2(ffma2(xa[i][1], u2, k2), make_float2(g0.z, g0.w), make_float2(b0.z, b0.w));
        xa[i][2] = ffma2(ffma2(xa[i][2], u2, k2), make_float2(g1.x, g1.y), make_float2(b1.x, b1.y));
        xa[i][3] = ffma2(ffma2(xa[i][3], u2, k2), make_float2(g1.z, g1.w), make_float2(b1.z, b1.w));
      }
    }
    if (pgt != nullptr) {
      const float2 u2 = make_float2(stt.y, stt.y), k2 = make_float2(-stt.x * stt.y, -stt.x * stt.y);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 g0 = vec[(2 * NV + i) * 64 + lane], g1 = vec[(2 * NV + i) * 64 + 32 + lane];
        const float4 b0 = vec[(3 * NV + i) * 64 + lane], b1 = vec[(3 * NV + i) * 64 + 32 + lane];
        xt[i][0] = ffma2(ffma2(xt[i][0], u2, k2), make_float2(g0.x, g0.y), make_float2(b0.x, b0.y));
        xt[i][1] = ffma2(ffma2(xt[i][1], u2, k2), make_float2(g0.z, g0.w), make_float2(b0.z, b0.w));
        xt[i][2] = ffma2(ffma2(xt[i][2], u2, k2), make_float2(g1.x, g1.y), make_float2(b1.x, b1.y));
        xt[i][3] = ffma2(ffma2(xt[i][3], u2, k2), make_float2(g1.z, g1.w), make_float2(b1.z, b1.w));
      }
    }
    float2 ra = make_float2(1.0f, 1.0f), ka = make_float2(0.0f, 0.0f), rt = ra, kt = ka;
    if (apply_ln) {
      float2 s1 = make_float2(0.0f, 0.0f), q1 = s1, s2 = s1, q2 = s1;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          s1 = fadd2(s1, xa[i][k]);
          q1 = ffma2(xa[i][k], xa[i][k], q1);
          s2 = fadd2(s2, xt[i][k]);
          q2 = ffma2(xt[i][k], xt[i][k], q2);
        }
      float v1 = s1.x + s1.y, v2 = q1.x + q1.y, v3 = s2.x + s2.y, v4 = q2.x + q2.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        v3 += __shfl_xor_sync(0xffffffffu, v3, o);
        v4 += __shfl_xor_sync(0xffffffffu, v4, o);
      }
      const float ma = v1 * inv_d, mt = v3 * inv_d;
      const float rsa = rsqrtf(fmaxf(v2 * inv_d - ma * ma, 0.0f) + eps), rst = rsqrtf(fmaxf(v4 * inv_d - mt * mt, 0.0f) + eps);
      ra = make_float2(rsa, rsa); ka = make_float2(-ma * rsa, -ma * rsa);
      rt = make_float2(rst, rst); kt = make_float2(-mt * rst, -mt * rst);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      float2 h[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        h[k] = ffma2(cA[i][k], ffma2(xa[i][k], ra, ka), ffma2(cT[i][k], ffma2(xt[i][k], rt, kt), cC[i][k]));
      if (c < d) {
        if (hb != nullptr) {
          uint4 o;
          o.x = pack_bf16(h[0].x, h[0].y); o.y = pack_bf16(h[1].x, h[1].y);
          o.z = pack_bf16(h[2].x, h[2].y); o.w = pack_bf16(h[3].x, h[3].y);
          *reinterpret_cast<uint4*>(hb + r * ldh + c) = o;
        }
        if (hf != nullptr) {
          float4* p = reinterpret_cast<float4*>(hf + r * ldh + c);
          p[0] = make_float4(h[0].x, h[0].y, h[1].x, h[1].y);
          p[1] = make_float4(h[2].x, h[2].y, h[3].x, h[3].y);
        }
      }
    }
    slot = slot + 1 == STAGES ? 0 : slot + 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ fused-LayerNorm helpers
// partials[slab][row] = (sum, sum of squares) over that slab's columns -> stats[row] = (mean, rstd)
__global__ void ln_stats_finalize_kernel(const float2* __restrict__ partials, int n_slabs, int64_t rows, int d,
                                         float eps, float2* __restrict__ stats) {
  const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (row >= rows) return;
  float s1 = 0.0f, s2 = 0.0f;
  for (int s = 0; s < n_slabs; ++s) {
    const float2 v = __ldg(partials + static_cast<int64_t>(s) * rows + row);
    s1 += v.x;
    s2 += v.y;
  }
  const float mean = s1 / static_cast<float>(d);
  const float var = fmaxf(s2 / static_cast<float>(d) - mean * mean, 0.0f);
  stats[row] = make_float2(mean, rsqrtf(var + eps));
}

// One warp per output row n of W [N,K]: W*gamma -> bf16, its row sum, and bias + W.beta
__global__ void fold_ln_weight_kernel(const float* __restrict__ W, int64_t ldw, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ bias,
                                      __nv_bfloat16* __restrict__ wf, int64_t ldo, float* __restrict__ colsum,
                                      float* __restrict__ bias_f, int N, int K) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  float cs = 0.0f, bs = 0.0f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[static_cast<int64_t>(n) * ldw + k];
    const __nv_bfloat16 h = __float2bfloat16_rn(w * gamma[k]);
    wf[static_cast<int64_t>(n) * ldo + k] = h;
    cs += __bfloat162float(h);
    bs = fmaf(w, beta[k], bs);
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    colsum[n] = cs;
    bias_f[n] = bs + (bias != nullptr ? bias[n] : 0.0f);
  }
}

__global__ void mean_over_time_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int L,
                                      int d) {
  const int64_t total = static_cast<int64_t>(B) * d;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = idx / d;
    const int c = static_cast<int>(idx - b * d);
    float s = 0.0f;
    for (int t = 0; t < L; ++t) s += x[(b * L + t) * d + c];
    out[idx] = s / static_cast<float>(L);
  }
}

// ------------------------------------------------------------------ backward of LayerNorm and ReLU (training)
// y = (x - mean) rstd gamma + beta.  dx = rstd (g - mean_c(g) - xhat mean_c(g xhat)), g = dy gamma; the parameter
// gradients dgamma = sum_rows dy xhat, dbeta = sum_rows dy are accumulated per warp in registers over a grid-stride
// loop, combined per CTA through shared memory and written as partial[cta][2][d] for a fixed-order final sum.
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_backward_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ dy,
                          int64_t lddy, const float* __restrict__ gamma, float eps, __nv_bfloat16* __restrict__ dx,
                          int64_t lddx, float* __restrict__ partial, int64_t rows, int d) {
  extern __shared__ float part[];   // [8 warps][2][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  RowRegs<NV> g_acc, b_acc, gm;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      g_acc.v[i][k] = 0.0f;
      b_acc.v[i][k] = 0.0f;
      gm.v[i][k] = c < d ? __ldg(gamma + c + k) : 0.0f;
    }
  }
  const float inv_d = 1.0f / static_cast<float>(d);
  for (int64_t row = blockIdx.x * 8 + warp; row < rows; row += static_cast<int64_t>(gridDim.x) * 8) {
    RowRegs<NV> xr, gr;
    load_row<false, NV>(xr, x + row * ldx, d, lane);
    load_row<false, NV>(gr, dy + row * lddy, d, lane);
    // ONE round of warp shuffles per row: sum x, sum x^2, sum g, sum g x (g = dy gamma) are reduced together, and
    // sum g xhat = rstd (sum g x - mean sum g).  (The two-pass statistics + two more sums cost four dependent rounds.)
    float sx = 0.0f, sxx = 0.0f, sg = 0.0f, sgx = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {   // columns past d hold zeros (load_row) and gm = 0 there
        const float xv = xr.v[i][k];
        const float g = gr.v[i][k] * gm.v[i][k];
        sx += xv;
        sxx = fmaf(xv, xv, sxx);
        sg += g;
        sgx = fmaf(g, xv, sgx);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sxx += __shfl_xor_sync(0xffffffffu, sxx, o);
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    const float mean = sx * inv_d;
    const float rstd = rsqrtf(fmaxf(sxx * inv_d - mean * mean, 0.0f) + eps);
    const float s1 = sg * inv_d;
    const float s2 = rstd * (sgx - mean * sg) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < d) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (xr.v[i][k] - mean) * rstd;
          const float dyv = gr.v[i][k];
          g_acc.v[i][k] = fmaf(dyv, xh, g_acc.v[i][k]);
          b_acc.v[i][k] += dyv;
          gr.v[i][k] = rstd * (dyv * gm.v[i][k] - s1 - xh * s2);
        }
      }
    }
    store_row(gr, dx + row * lddx, nullptr, d, lane);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        part[(warp * 2 + 0) * d + c + k] = g_acc.v[i][k];
        part[(warp * 2 + 1) * d + c + k] = b_acc.v[i][k];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
    const int which = c / d, col = c - which * d;
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[(w * 2 + which) * d + col];
    partial[static_cast<int64_t>(blockIdx.x) * 2 * d + c] = s;
  }
}

// The same, streaming form (default): one persistent CTA per SM, rows reach a warp through a private cp.async ring of
// (x | dy) row pairs (STAGES - 1 pairs = 15 KB in flight per warp at d = 768; the kernel above has one pair per warp in
// flight and nothing to overlap its ~600 instructions per row with: 0.5 of the HBM roofline), arithmetic in packed fp32.
template <int NV>
__global__ void __launch_bounds__(256, 1)
layernorm_backward_ring_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ dy,
                               int64_t lddy, const float* __restrict__ gamma, float eps, __nv_bfloat16* __restrict__ dx,
                               int64_t lddx, float* __restrict__ partial, int64_t rows, int d) {
  constexpr int STAGES = GbStages<NV>::value;
  constexpr int PAIR_BYTES = 2 * NV * 512;
  extern __shared__ __align__(16) uint8_t lnb_smem[];   // ring [8 warps][STAGES][x | dy][NV * 32 lanes][16 B]; later part[8][2][d]
  float* part = reinterpret_cast<float*>(lnb_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring_p = lnb_smem + warp * (STAGES * PAIR_BYTES) + lane * 16;
  const uint32_t ring = smem_u32(ring_p);
  float2 g_acc[NV][4], b_acc[NV][4], gm[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      g_acc[i][k] = b_acc[i][k] = make_float2(0.0f, 0.0f);
      gm[i][k] = c < d ? make_float2(__ldg(gamma + c + 2 * k), __ldg(gamma + c + 2 * k + 1)) : make_float2(0.0f, 0.0f);
    }
  }
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 8;
  auto issue = [&](int64_t row, int slot) {   // one commit group per call, also when there is nothing left to copy
    if (row < rows) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 8;
        if (c < d) {
          lmm_cp_async16(ring + slot * PAIR_BYTES + i * 512, x + row * ldx + c);
          lmm_cp_async16(ring + slot * PAIR_BYTES + (NV + i) * 512, dy + row * lddy + c);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * 8 + warp;
#pragma unroll
  for (int s0 = 0; s0 < STAGES - 1; ++s0) issue(row0 + s0 * stride, s0);
  const float inv_d = 1.0f / static_cast<float>(d);
  int slot = 0;
  for (int64_t row = row0; row < rows; row += stride) {
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
    float2 xv[NV][4], gv[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      uint4 u = make_uint4(0u, 0u, 0u, 0u), v = make_uint4(0u, 0u, 0u, 0u);
      if (c < d) {
        u = *reinterpret_cast<const uint4*>(ring_p + slot * PAIR_BYTES + i * 512);
        v = *reinterpret_cast<const uint4*>(ring_p + slot * PAIR_BYTES + (NV + i) * 512);
      }
      xv[i][0] = make_float2(bf16_lo(u.x), bf16_hi(u.x)); xv[i][1] = make_float2(bf16_lo(u.y), bf16_hi(u.y));
      xv[i][2] = make_float2(bf16_lo(u.z), bf16_hi(u.z)); xv[i][3] = make_float2(bf16_lo(u.w), bf16_hi(u.w));
      gv[i][0] = make_float2(bf16_lo(v.x), bf16_hi(v.x)); gv[i][1] = make_float2(bf16_lo(v.y), bf16_hi(v.y));
      gv[i][2] = make_float2(bf16_lo(v.z), bf16_hi(v.z)); gv[i][3] = make_float2(bf16_lo(v.w), bf16_hi(v.w));
    }
    // the slot consumed in the PREVIOUS iteration takes the row STAGES - 1 strides ahead
    issue(row + (STAGES - 1) * stride, slot == 0 ? STAGES - 1 : slot - 1);
    // sum x, sum x^2, sum g, sum g x (g = dy gamma) reduced together; columns past d hold zeros and gm = 0 there
    float2 sx = make_float2(0.0f, 0.0f), sxx = sx, sg = sx, sgx = sx;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 g = fmul2(gv[i][k], gm[i][k]);
        sx = fadd2(sx, xv[i][k]);
        sxx = ffma2(xv[i][k], xv[i][k], sxx);
        sg = fadd2(sg, g);
        sgx = ffma2(g, xv[i][k], sgx);
      }
    float v1 = sx.x + sx.y, v2 = sxx.x + sxx.y, v3 = sg.x + sg.y, v4 = sgx.x + sgx.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o);
      v3 += __shfl_xor_sync(0xffffffffu, v3, o);
      v4 += __shfl_xor_sync(0xffffffffu, v4, o);
    }
    const float mean = v1 * inv_d;
    const float rstd = rsqrtf(fmaxf(v2 * inv_d - mean * mean, 0.0f) + eps);
    const float s1 = v3 * inv_d;
    const float s2 = rstd * (v4 - mean * v3) * inv_d;
    const float2 r2 = make_float2(rstd, rstd), mr = make_float2(-mean * rstd, -mean * rstd);
    const float2 a1 = make_float2(-s1 * rstd, -s1 * rstd), a2 = make_float2(-s2 * rstd, -s2 * rstd);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      float2 o2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 xh = ffma2(xv[i][k], r2, mr);
        g_acc[i][k] = ffma2(gv[i][k], xh, g_acc[i][k]);
        b_acc[i][k] = fadd2(b_acc[i][k], gv[i][k]);
        // rstd (dy gamma - s1 - xhat s2)
        o2[k] = ffma2(xh, a2, ffma2(fmul2(gv[i][k], gm[i][k]), r2, a1));
      }
      if (c < d) {
        uint4 o;
        o.x = pack_bf16(o2[0].x, o2[0].y); o.y = pack_bf16(o2[1].x, o2[1].y);
        o.z = pack_bf16(o2[2].x, o2[2].y); o.w = pack_bf16(o2[3].x, o2[3].y);
        *reinterpret_cast<uint4*>(dx + row * lddx + c) = o;
      }
    }
    slot = slot + 1 == STAGES ? 0 : slot + 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();   // every warp is done with its ring: the memory becomes part[8][2][d]
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        part[(warp * 2 + 0) * d + c + 2 * k] = g_acc[i][k].x;
        part[(warp * 2 + 0) * d + c + 2 * k + 1] = g_acc[i][k].y;
        part[(warp * 2 + 1) * d + c + 2 * k] = b_acc[i][k].x;
        part[(warp * 2 + 1) * d + c + 2 * k + 1] = b_acc[i][k].y;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
    const int which = c / d, col = c - which * d;
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[(w * 2 + which) * d + col];
    partial[static_cast<int64_t>(blockIdx.x) * 2 * d + c] = s;
  }
}

// out[0..d) = dgamma, out2[0..d) = dbeta: fixed-order sum of the per-CTA partials (optionally on top of the old value)
__global__ void __launch_bounds__(256)
ln_param_grad_finalize_kernel(const float* __restrict__ partial, int n_ctas, int d, float* __restrict__ dgamma,
                              float* __restrict__ dbeta, int accumulate) {
  // 32 columns per CTA; the 8 warps each sum every 8th partial (coalesced 128-byte reads), then a fixed-order sum of the
  // 8 warp sums through shared memory: deterministic, and 1 / 8 of the serial chain of the thread-per-column loop
  __shared__ float wsum[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  float s = 0.0f;
  if (c < 2 * d)
    for (int i = warp; i < n_ctas; i += 8) s += partial[static_cast<int64_t>(i) * 2 * d + c];
  wsum[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < 2 * d) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += wsum[w][lane];
    float* dst = c < d ? dgamma + c : dbeta + (c - d);
    *dst = (accumulate ? *dst : 0.0f) + t;
  }
}

// dx = dy where h > 0 else 0 (h = the ReLU's OUTPUT, which is what the forward keeps): 8 elements per thread
__global__ void relu_backward_kernel(const __nv_bfloat16* __restrict__ dy, int64_t lddy, const __nv_bfloat16* __restrict__ h,
                                     int64_t ldh, __nv_bfloat16* __restrict__ dx, int64_t lddx, int64_t rows, int cols) {
  const int64_t chunks = cols / 8, total = rows * chunks;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = idx / chunks;
    const int c = static_cast<int>(idx - r * chunks) * 8;
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(dy + r * lddy + c));
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(h + r * ldh + c));
    uint4 o;
    // a bf16 is positive iff its sign bit is clear and it is not zero
    auto keep = [](uint32_t gv, uint32_t av) {
      const uint32_t lo = ((av & 0x8000u) == 0u && (av & 0x7fffu) != 0u) ? (gv & 0xffffu) : 0u;
      const uint32_t hi = ((av & 0x80000000u) == 0u && (av & 0x7fff0000u) != 0u) ? (gv & 0xffff0000u) : 0u;
      return lo | hi;
    };
    o.x = keep(g.x, a.x); o.y = keep(g.y, a.y); o.z = keep(g.z, a.z); o.w = keep(g.w, a.w);
    *reinterpret_cast<uint4*>(dx + r * lddx + c) = o;
  }
}

static unsigned grid_for(int64_t work_items, int block) {
  int64_t g = (work_items + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

// NV dispatch: instantiate the row kernels for 1, 2, 3, 4 and 8 vectors per lane (d <= 256 ... 2048)
#define HRIEMO_DISPATCH_NV(d, ...)                              \
  do {                                                          \
    const int nv_ = ((d) + 255) / 256;                          \
    if (nv_ <= 1) { constexpr int NV = 1; __VA_ARGS__; }        \
    else if (nv_ == 2) { constexpr int NV = 2; __VA_ARGS__; }   \
    else if (nv_ == 3) { constexpr int NV = 3; __VA_ARGS__; }   \
    else if (nv_ == 4) { constexpr int NV = 4; __VA_ARGS__; }   \
    else { constexpr int NV = 8; __VA_ARGS__; }                 \
  } while (0)

static bool row_shape_ok(int d) { return d > 0 && d % 8 == 0 && d <= ROW_MAX_VEC * 256; }
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace hriemo

using namespace hriemo;

extern "C" int hriemo_cast_f32_to_bf16(const float* in, int64_t ld_in, void* out, int64_t ld_out,
                                       int64_t rows, int32_t cols, void* stream) {
  HRIEMO_REQUIRE(in && out, "cast: null pointer");
  HRIEMO_REQUIRE(rows >= 0 && cols > 0 && ld_in >= cols && ld_out >= cols && ld_out % 8 == 0,
                 "cast: bad shape rows=%lld cols=%d ld_in=%lld ld_out=%lld", (long long)rows, cols,
                 (long long)ld_in, (long long)ld_out);
  HRIEMO_REQUIRE(aligned16(out), "cast: out misaligned");
  if (rows == 0) return HRIEMO_OK;
  const int vec_ok = aligned16(in) && ld_in % 4 == 0;
  const int64_t total = rows * (ld_out / 8);
  cast_f32_bf16_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, ld_in, static_cast<__nv_bfloat16*>(out), ld_out, rows, cols, vec_ok);
  return check_launch("cast_f32_to_bf16");
}

extern "C" int hriemo_layernorm(const void* x, int32_t x_is_f32, int64_t ldx, const float* gamma,
                                const float* beta, float eps, void* y_bf16, float* y_f32, int64_t ldy,
                                int64_t rows, int32_t d, void* stream) {
  HRIEMO_REQUIRE(x && gamma && beta && (y_bf16 || y_f32), "layernorm: null pointer");
  HRIEMO_REQUIRE(row_shape_ok(d), "layernorm: d=%d must be a multiple of 8 and <= %d", d, ROW_MAX_VEC * 256);
  HRIEMO_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(gamma) && aligned16(beta) &&
                     aligned16(y_bf16) && aligned16(y_f32),
                 "layernorm: misaligned operand");
  if (rows <= 0) return HRIEMO_OK;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y_bf16);
  if (x_is_f32)
    HRIEMO_DISPATCH_NV(d, (layernorm_kernel<true, NV><<<grid, 256, 0, s>>>(x, ldx, gamma, beta, eps, yb, y_f32, ldy, rows, d)));
  else
    HRIEMO_DISPATCH_NV(d, (layernorm_kernel<false, NV><<<grid, 256, 0, s>>>(x, ldx, gamma, beta, eps, yb, y_f32, ldy, rows, d)));
  return check_launch("layernorm");
}

extern "C" int hriemo_ln_masked_mean(const void* x, int64_t ldx, const float* gamma, const float* beta,
                                     float eps, int32_t apply_ln, const uint8_t* pad, float* pooled,
                                     int64_t ld_pooled, int32_t B, int32_t T, int32_t d, const float* pre_gamma,
                                     const float* pre_beta, const float* pre_stats, void* stream) {
  HRIEMO_REQUIRE(x && pooled && (!apply_ln || (gamma && beta)), "ln_masked_mean: null pointer");
  HRIEMO_REQUIRE((pre_gamma == nullptr) == (pre_beta == nullptr) && aligned16(pre_gamma) && aligned16(pre_beta),
                 "ln_masked_mean: pre_gamma / pre_beta go together, 16-byte aligned");
  HRIEMO_REQUIRE(row_shape_ok(d) && B > 0 && T > 0, "ln_masked_mean: bad shape B=%d T=%d d=%d", B, T, d);
  HRIEMO_REQUIRE(ldx % 8 == 0 && aligned16(x) && aligned16(gamma) && aligned16(beta),
                 "ln_masked_mean: misaligned operand");
  // warp-private row rings (8 warps x STAGES x NV * 512 bytes; they hold the 8 x d partial sums afterwards): opt-in size
  static uint64_t lmm_attr_done = 0;
  if (device_needs_attr(&lmm_attr_done)) {
    cudaError_t e = cudaSuccess;
#define HRIEMO_LMM_ATTR(NVV)                                                                                  \
    if (e == cudaSuccess)                                                                                     \
      e = cudaFuncSetAttribute(ln_masked_mean_kernel<NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                               LMM_WARPS * LmmStages<NVV>::value * NVV * 512)
    HRIEMO_LMM_ATTR(1); HRIEMO_LMM_ATTR(2); HRIEMO_LMM_ATTR(3); HRIEMO_LMM_ATTR(4); HRIEMO_LMM_ATTR(8);
#undef HRIEMO_LMM_ATTR
    if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "ln_masked_mean: %s", cudaGetErrorString(e));
  }
  HRIEMO_DISPATCH_NV(d, (ln_masked_mean_kernel<NV><<<B, 256, LMM_WARPS * LmmStages<NV>::value * NV * 512, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, eps, apply_ln, pad, pooled, ld_pooled, T, d,
      pre_gamma, pre_beta, reinterpret_cast<const float2*>(pre_stats))));
  return check_launch("ln_masked_mean");
}

extern "C" int hriemo_gate_input(const float* a_pool, const float* t_pool, float* g, int32_t B, int32_t d,
                                 void* stream) {
  HRIEMO_REQUIRE(a_pool && t_pool && g && B > 0 && d > 0, "gate_input: bad argument");
  gate_input_kernel<<<grid_for(static_cast<int64_t>(B) * d, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a_pool, t_pool, g, B, d);
  return check_launch("gate_input");
}

extern "C" int hriemo_gate_blend(const void* a, int64_t lda, int32_t T_a, const void* t, int64_t ldt,
                                 const float* gamma_a, const float* beta_a, const float* gamma_t,
                                 const float* beta_t, float eps, int32_t apply_ln, const float* w,
                                 int32_t w_is_scalar, void* h_bf16, float* h_f32, int64_t ldh,
                                 float* beta_out, int32_t B, int32_t L, int32_t d, const float* pre_gamma_a,
                                 const float* pre_beta_a, const float* pre_gamma_t, const float* pre_beta_t,
                                 const float* pre_stats_a, const float* pre_stats_t, void* stream) {
  HRIEMO_REQUIRE(a && t && w && (h_bf16 || h_f32), "gate_blend: null pointer");
  HRIEMO_REQUIRE((pre_gamma_a == nullptr) == (pre_beta_a == nullptr) && (pre_gamma_t == nullptr) == (pre_beta_t == nullptr) &&
                     aligned16(pre_gamma_a) && aligned16(pre_beta_a) && aligned16(pre_gamma_t) && aligned16(pre_beta_t),
                 "gate_blend: pre-LayerNorm parameters come in (gamma, beta) pairs, 16-byte aligned");
  HRIEMO_REQUIRE(!apply_ln || (gamma_a && beta_a && gamma_t && beta_t), "gate_blend: LN params missing");
  HRIEMO_REQUIRE(row_shape_ok(d) && B > 0 && L > 0 && T_a >= L, "gate_blend: bad shape B=%d L=%d T_a=%d d=%d",
                 B, L, T_a, d);
  HRIEMO_REQUIRE(lda % 8 == 0 && ldt % 8 == 0 && ldh % 8 == 0 && aligned16(a) && aligned16(t) &&
                     aligned16(h_bf16) && aligned16(h_f32),
                 "gate_blend: misaligned operand");
  const int64_t rows = static_cast<int64_t>(B) * L;
  static const bool force_v1 = getenv("HRIEMO_GATE_BLEND_V1") != nullptr;   // the warp-per-row form, for A / B runs
  if (!force_v1 && d <= 1024 /* wider rows spill in the streaming form */ && (pre_gamma_a == nullptr || pre_stats_a != nullptr) && (pre_gamma_t == nullptr || pre_stats_t != nullptr) &&
      (w_is_scalar || aligned16(w)) && (!apply_ln || (aligned16(gamma_a) && aligned16(beta_a) && aligned16(gamma_t) && aligned16(beta_t)))) {
    // streaming form: persistent CTAs, warp-private row rings (gate_blend_stream_kernel)
    static uint64_t gb_attr_done = 0;
    if (device_needs_attr(&gb_attr_done)) {
      cudaError_t e = cudaSuccess;
#define HRIEMO_GB_ATTR(NVV)                                                                                      \
      if (e == cudaSuccess)                                                                                      \
        e = cudaFuncSetAttribute(gate_blend_stream_kernel<NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                 4 * NVV * 1024 + 8 * GbStages<NVV>::value * 2 * NVV * 512)
      HRIEMO_GB_ATTR(1); HRIEMO_GB_ATTR(2); HRIEMO_GB_ATTR(3); HRIEMO_GB_ATTR(4); HRIEMO_GB_ATTR(8);
#undef HRIEMO_GB_ATTR
      if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "gate_blend: %s", cudaGetErrorString(e));
    }
    int64_t grid = (rows + 7) / 8;
    const int64_t resident = static_cast<int64_t>(device_sm_count()) * (d <= 256 ? 2 : 1);   // NV = 1: two CTAs fit an SM (118 registers, 53 KB)
    if (grid > resident) grid = resident;
    HRIEMO_DISPATCH_NV(d, (gate_blend_stream_kernel<NV><<<static_cast<unsigned>(grid), 256, 4 * NV * 1024 + 8 * GbStages<NV>::value * 2 * NV * 512,
                                                         static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(a), lda, T_a, static_cast<const __nv_bfloat16*>(t), ldt, gamma_a,
        beta_a, gamma_t, beta_t, eps, apply_ln, w, w_is_scalar, static_cast<__nv_bfloat16*>(h_bf16), h_f32,
        ldh, beta_out, B, L, d, pre_gamma_a, pre_beta_a, pre_gamma_t, pre_beta_t,
        reinterpret_cast<const float2*>(pre_stats_a), reinterpret_cast<const float2*>(pre_stats_t))));
    return check_launch("gate_blend");
  }
  HRIEMO_DISPATCH_NV(d, (gate_blend_kernel<NV><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), lda, T_a, static_cast<const __nv_bfloat16*>(t), ldt, gamma_a,
      beta_a, gamma_t, beta_t, eps, apply_ln, w, w_is_scalar, static_cast<__nv_bfloat16*>(h_bf16), h_f32,
      ldh, beta_out, B, L, d, pre_gamma_a, pre_beta_a, pre_gamma_t, pre_beta_t,
      reinterpret_cast<const float2*>(pre_stats_a), reinterpret_cast<const float2*>(pre_stats_t))));
  return check_launch("gate_blend");
}

extern "C" int hriemo_ln_stats_finalize(const float* partials, int32_t n_slabs, int64_t rows, int32_t d, float eps,
                                        float* stats, void* stream) {
  HRIEMO_REQUIRE(partials && stats && n_slabs > 0 && rows >= 0 && d > 0, "ln_stats_finalize: bad argument");
  HRIEMO_REQUIRE(((reinterpret_cast<uintptr_t>(partials) | reinterpret_cast<uintptr_t>(stats)) & 7u) == 0,
                 "ln_stats_finalize: misaligned operand");
  if (rows == 0) return HRIEMO_OK;
  ln_stats_finalize_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(partials), n_slabs, rows, d, eps, reinterpret_cast<float2*>(stats));
  return check_launch("ln_stats_finalize");
}

extern "C" int hriemo_fold_ln_weight(const float* W, int64_t ldw, const float* gamma, const float* beta,
                                     const float* bias, void* w_folded, int64_t ldo, float* colsum,
                                     float* bias_folded, int32_t N, int32_t K, void* stream) {
  HRIEMO_REQUIRE(W && gamma && beta && w_folded && colsum && bias_folded, "fold_ln_weight: null pointer");
  HRIEMO_REQUIRE(N > 0 && K > 0 && ldw >= K && ldo >= K, "fold_ln_weight: bad shape N=%d K=%d", N, K);
  fold_ln_weight_kernel<<<(N + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, ldw, gamma, beta, bias, static_cast<__nv_bfloat16*>(w_folded), ldo, colsum, bias_folded, N, K);
  return check_launch("fold_ln_weight");
}

extern "C" int hriemo_mean_over_time(const float* x, float* out, int32_t B, int32_t L, int32_t d,
                                     void* stream) {
  HRIEMO_REQUIRE(x && out && B > 0 && L > 0 && d > 0, "mean_over_time: bad argument");
  mean_over_time_kernel<<<grid_for(static_cast<int64_t>(B) * d, 256), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(x, out, B, L, d);
  return check_launch("mean_over_time");
}

static int ln_backward_ctas(int64_t rows) {
  int64_t g = (rows + 7) / 8;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 2;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

extern "C" int64_t hriemo_layernorm_backward_workspace_bytes(int64_t rows, int32_t d) {
  return static_cast<int64_t>(ln_backward_ctas(rows)) * 2 * d * static_cast<int64_t>(sizeof(float));
}

extern "C" int hriemo_layernorm_backward(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* gamma,
                                         float eps, void* dx, int64_t lddx, float* dgamma, float* dbeta,
                                         int32_t accumulate, void* workspace, int64_t rows, int32_t d, void* stream) {
  HRIEMO_REQUIRE(x && dy && gamma && dx && dgamma && dbeta && workspace, "layernorm_backward: null pointer");
  HRIEMO_REQUIRE(row_shape_ok(d) && d <= 1024 && rows > 0, "layernorm_backward: d=%d must be a multiple of 8 and <= 1024", d);
  HRIEMO_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx) &&
                     aligned16(gamma),
                 "layernorm_backward: misaligned operand");
  const int ctas = ln_backward_ctas(rows);
  const size_t smem = static_cast<size_t>(8) * 2 * d * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  const __nv_bfloat16 *xp = static_cast<const __nv_bfloat16*>(x), *gp = static_cast<const __nv_bfloat16*>(dy);
  __nv_bfloat16* dp = static_cast<__nv_bfloat16*>(dx);
  const int nv = (d + 255) / 256;
  static const bool force_v1 = getenv("HRIEMO_LN_BWD_V1") != nullptr;   // the one-row-per-warp form, for A / B runs
  if (!force_v1) {
    // streaming form: one persistent CTA per SM, warp-private (x | dy) rings (layernorm_backward_ring_kernel)
    static uint64_t lnb_attr_done = 0;
    if (device_needs_attr(&lnb_attr_done)) {
      cudaError_t e = cudaSuccess;
#define HRIEMO_LNB_ATTR(NVV)                                                                                         \
      if (e == cudaSuccess)                                                                                          \
        e = cudaFuncSetAttribute(layernorm_backward_ring_kernel<NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 8 * GbStages<NVV>::value * 2 * NVV * 512)
      HRIEMO_LNB_ATTR(1); HRIEMO_LNB_ATTR(2); HRIEMO_LNB_ATTR(3); HRIEMO_LNB_ATTR(4);
#undef HRIEMO_LNB_ATTR
      if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "layernorm_backward: %s", cudaGetErrorString(e));
    }
    const int rctas = ctas < device_sm_count() ? ctas : device_sm_count();   // the workspace holds ln_backward_ctas(rows) >= rctas partials
#define HRIEMO_LNB_LAUNCH(NVV)                                                                                       \
    layernorm_backward_ring_kernel<NVV><<<rctas, 256, 8 * GbStages<NVV>::value * 2 * NVV * 512, s>>>(xp, ldx, gp, lddy, gamma, eps, dp, \
                                                                                                     lddx, partial, rows, d)
    if (nv <= 1) HRIEMO_LNB_LAUNCH(1); else if (nv == 2) HRIEMO_LNB_LAUNCH(2); else if (nv == 3) HRIEMO_LNB_LAUNCH(3); else HRIEMO_LNB_LAUNCH(4);
#undef HRIEMO_LNB_LAUNCH
    int rc = check_launch("layernorm_backward");
    if (rc) return rc;
    ln_param_grad_finalize_kernel<<<(2 * d + 31) / 32, 256, 0, s>>>(partial, rctas, d, dgamma, dbeta, accumulate);
    return check_launch("layernorm_backward (parameter gradients)");
  }
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(layernorm_backward_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4);
  }
  if (nv <= 1) layernorm_backward_kernel<1><<<ctas, 256, smem, s>>>(xp, ldx, gp, lddy, gamma, eps, dp, lddx, partial, rows, d);
  else if (nv == 2) layernorm_backward_kernel<2><<<ctas, 256, smem, s>>>(xp, ldx, gp, lddy, gamma, eps, dp, lddx, partial, rows, d);
  else if (nv == 3) layernorm_backward_kernel<3><<<ctas, 256, smem, s>>>(xp, ldx, gp, lddy, gamma, eps, dp, lddx, partial, rows, d);
  else layernorm_backward_kernel<4><<<ctas, 256, smem, s>>>(xp, ldx, gp, lddy, gamma, eps, dp, lddx, partial, rows, d);
  int rc = check_launch("layernorm_backward");
  if (rc) return rc;
  ln_param_grad_finalize_kernel<<<(2 * d + 31) / 32, 256, 0, s>>>(partial, ctas, d, dgamma, dbeta, accumulate);
  return check_launch("layernorm_backward (parameter gradients)");
}

extern "C" int hriemo_relu_backward_bf16(const void* dy, int64_t lddy, const void* h, int64_t ldh, void* dx, int64_t lddx,
                                         int64_t rows, int32_t cols, void* stream) {
  HRIEMO_REQUIRE(dy && h && dx && rows >= 0 && cols > 0 && cols % 8 == 0, "relu_backward: cols must be a multiple of 8");
  HRIEMO_REQUIRE(lddy % 8 == 0 && ldh % 8 == 0 && lddx % 8 == 0 && aligned16(dy) && aligned16(h) && aligned16(dx),
                 "relu_backward: misaligned operand");
  if (rows == 0) return HRIEMO_OK;
  relu_backward_kernel<<<grid_for(rows * (cols / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), lddy, static_cast<const __nv_bfloat16*>(h), ldh, static_cast<__nv_bfloat16*>(dx),
      lddx, rows, cols);
  return check_launch("relu_backward");
}
