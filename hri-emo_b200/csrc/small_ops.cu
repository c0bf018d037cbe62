// CUDA-core kernels for the small, decision-sensitive parts of the path:
//  * fp32 GEMM for the beta-gate MLP, the emotion head and the classifier head,
//  * small-query attention (emotion decoder) and head-averaged attention maps.
#include <math.h>
#include <stdlib.h>

#include "dropout.cuh"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

// ------------------------------------------------------------------ fp32 GEMM
// out[M,N] = act(A[M,K] . W[N,K]^T + bias); 64x64 tile, 16-wide K slab, 4x4 per thread.
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__global__ void __launch_bounds__(256)
sgemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
                 const float* __restrict__ bias, float* __restrict__ out, int64_t ldo, int64_t M, int N,
                 int K, int act) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * SG_BM;
  const int n0 = blockIdx.x * SG_BN;
  const bool relu_in = (act & 4) != 0;   // ReLU applied to A as it is read (the hidden layer arrives pre-activation)
  act &= 3;
  float acc[4][4] = {};
  const int lr = tid >> 2;        // 0..63: row inside the tile
  const int lk = (tid & 3) * 4;   // 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lk + i;
      const int64_t am = m0 + lr;
      const float av = (am < M && k < K) ? __ldg(A + am * lda + k) : 0.0f;
      As[lk + i][lr] = (relu_in && av < 0.0f) ? 0.0f : av;   // relu that keeps NaN (fmaxf would turn it into 0; torch.relu does not)
      const int wn = n0 + lr;
      Ws[lk + i][lr] = (wn < N && k < K) ? __ldg(W + static_cast<int64_t>(wn) * ldw + k) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.0f);
      if (act == 1) v = v < 0.0f ? 0.0f : v;   // NaN stays NaN, as in torch.relu: the gate of a fully padded utterance must come out NaN
      else if (act == 2) v = 1.0f / (1.0f + expf(-v));
      out[m * ldo + n] = v;
    }
  }
}

// ------------------------------------------------------------------ fp32 GEMM with a handful of output columns
// out[M, N <= 4] = act(A . W^T + bias): the emotion head (N = 1, one logit per query row).  The 64 x 64 tile kernel above
// spends 63 / 64 of its work on columns that do not exist (59 us for 8192 x 768 rows); here a warp owns a row, reads it once
// with 16-byte loads and holds the N dot products in registers (same fp32 FMA arithmetic, a different summation order).
constexpr int SGN_MAX_N = 4;

__global__ void __launch_bounds__(256)
sgemm_few_cols_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
                      const float* __restrict__ bias, float* __restrict__ out, int64_t ldo, int64_t M, int N, int K, int act) {
  const bool relu_in = (act & 4) != 0;
  act &= 3;
  const int lane = threadIdx.x & 31;
  for (int64_t m = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); m < M;
       m += static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5)) {
    float acc[SGN_MAX_N];
#pragma unroll
    for (int n = 0; n < SGN_MAX_N; ++n) acc[n] = 0.0f;
    const float* a = A + m * lda;
    for (int k = lane * 4; k < K; k += 128) {
      float4 av = __ldg(reinterpret_cast<const float4*>(a + k));
      if (relu_in) {   // relu that keeps NaN, as torch.relu does
        av.x = av.x < 0.0f ? 0.0f : av.x; av.y = av.y < 0.0f ? 0.0f : av.y;
        av.z = av.z < 0.0f ? 0.0f : av.z; av.w = av.w < 0.0f ? 0.0f : av.w;
      }
#pragma unroll
      for (int n = 0; n < SGN_MAX_N; ++n) {
        if (n < N) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(n) * ldw + k));
          acc[n] = fmaf(av.x, wv.x, fmaf(av.y, wv.y, fmaf(av.z, wv.z, fmaf(av.w, wv.w, acc[n]))));
        }
      }
    }
#pragma unroll
    for (int n = 0; n < SGN_MAX_N; ++n) {
      if (n < N) {
        float v = acc[n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
          v += bias ? __ldg(bias + n) : 0.0f;
          if (act == 1) v = v < 0.0f ? 0.0f : v;
          else if (act == 2) v = 1.0f / (1.0f + expf(-v));
          out[m * ldo + n] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ small attention
// One CTA per (utterance, block of up to 8 query rows); loops over heads.
// Per head: scores for every key (thread per key), exact softmax per query row
// (warp per row), optional head-averaged probabilities, then P.V (thread per column).
constexpr int SA_NQ = 8;
constexpr int SA_THREADS = 128;

__global__ void __launch_bounds__(SA_THREADS)
small_attention_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k,
                       int64_t ldk, const __nv_bfloat16* __restrict__ v, int64_t ldv,
                       const uint8_t* __restrict__ key_pad, __nv_bfloat16* __restrict__ out, int64_t ldo,
                       float* __restrict__ probs, int H, int Nq, int Tk, int dh, float scale, uint32_t drop_p8,
                       uint32_t drop_key, float drop_scale) {
  extern __shared__ float sm[];
  float* qs = sm;                         // [SA_NQ][dh]
  float* sc = qs + SA_NQ * dh;            // [SA_NQ][Tk]  scores -> probabilities
  float* pavg = sc + SA_NQ * Tk;          // [SA_NQ][Tk]  (only if probs)
  const int b = blockIdx.x;
  const int qb = blockIdx.y * SA_NQ;
  const int nq = min(SA_NQ, Nq - qb);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float inv_h = 1.0f / static_cast<float>(H);
  if (probs)
    for (int i = tid; i < SA_NQ * Tk; i += SA_THREADS) pavg[i] = 0.0f;

  // gridDim.z == H: one head per CTA (no head-averaged map to accumulate); else this CTA loops over heads
  const int h_begin = gridDim.z > 1 ? blockIdx.z : 0;
  const int h_end = gridDim.z > 1 ? blockIdx.z + 1 : H;
  for (int h = h_begin; h < h_end; ++h) {
    __syncthreads();
    for (int i = tid; i < nq * dh; i += SA_THREADS) {
      const int qi = i / dh, c = i - qi * dh;
      qs[qi * dh + c] =
          __bfloat162float(q[(static_cast<int64_t>(b) * Nq + qb + qi) * ldq + h * dh + c]) * scale;
    }
    __syncthreads();
    // scores
    for (int j = tid; j < Tk; j += SA_THREADS) {
      const bool pad = key_pad != nullptr && key_pad[static_cast<int64_t>(b) * Tk + j] != 0;
      float acc[SA_NQ];
#pragma unroll
      for (int qi = 0; qi < SA_NQ; ++qi) acc[qi] = 0.0f;
      const __nv_bfloat16* kr = k + (static_cast<int64_t>(b) * Tk + j) * ldk + h * dh;
      for (int c = 0; c < dh; c += 8) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(kr + c));
        const float kv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                             bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
        for (int qi = 0; qi < SA_NQ; ++qi) {
          if (qi < nq) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[qi] = fmaf(qs[qi * dh + c + e], kv[e], acc[qi]);
          }
        }
      }
#pragma unroll
      for (int qi = 0; qi < SA_NQ; ++qi)
        if (qi < nq) sc[qi * Tk + j] = pad ? -INFINITY : acc[qi];
    }
    __syncthreads();
    // softmax per query row (a fully masked row gives NaN, like torch.softmax)
    for (int qi = warp; qi < nq; qi += SA_THREADS / 32) {
      float m = -INFINITY;
      for (int j = lane; j < Tk; j += 32) m = fmaxf(m, sc[qi * Tk + j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s = 0.0f;
      for (int j = lane; j < Tk; j += 32) {
        const float e = expf(sc[qi * Tk + j] - m);
        sc[qi * Tk + j] = e;
        s += e;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float inv = 1.0f / s;
      // training: dropout on the probabilities after the normalisation (csrc/dropout.cuh; the map, if asked for, is
      // the undropped one)
      const uint32_t dkey = drop_key_bh(drop_key, static_cast<uint32_t>(b * H + h));
      for (int j = lane; j < Tk; j += 32) {
        const float pr = sc[qi * Tk + j] * inv;
        if (probs) pavg[qi * Tk + j] += pr * inv_h;
        sc[qi * Tk + j] = (drop_p8 == 0u || drop_keep(dkey, static_cast<uint32_t>(qb + qi), static_cast<uint32_t>(j), drop_p8))
                              ? pr * drop_scale : 0.0f;
      }
    }
    __syncthreads();
    // P.V
    if (out != nullptr) {
      for (int c = tid; c < dh; c += SA_THREADS) {
        float acc[SA_NQ];
#pragma unroll
        for (int qi = 0; qi < SA_NQ; ++qi) acc[qi] = 0.0f;
        const __nv_bfloat16* vc = v + static_cast<int64_t>(b) * Tk * ldv + h * dh + c;
        for (int j = 0; j < Tk; ++j) {
          const float vv = __bfloat162float(vc[static_cast<int64_t>(j) * ldv]);
#pragma unroll
          for (int qi = 0; qi < SA_NQ; ++qi)
            if (qi < nq) acc[qi] = fmaf(sc[qi * Tk + j], vv, acc[qi]);
        }
#pragma unroll
        for (int qi = 0; qi < SA_NQ; ++qi)
          if (qi < nq)
            out[(static_cast<int64_t>(b) * Nq + qb + qi) * ldo + h * dh + c] = __float2bfloat16_rn(acc[qi]);
      }
    }
  }
  if (probs) {
    __syncthreads();
    for (int i = tid; i < nq * Tk; i += SA_THREADS) {
      const int qi = i / Tk, j = i - qi * Tk;
      probs[(static_cast<int64_t>(b) * Nq + qb + qi) * Tk + j] = pavg[qi * Tk + j];
    }
  }
}

// ------------------------------------------------------------------ decoder attention (the common case of the above)
// N_q <= 8 learned queries against T_k <= 128 keys, no attention map requested (models/emotion_decoder.py:42, :48-54 with
// return_attention False): the work is reading the memory's [K|V] projection once.  One CTA per (utterance, head); the
// head's K and V tiles ([T_k][dh] bf16, 12 KB each at 64 x 96) are brought into shared memory with 16-byte cp.async
// copies issued back to back (24 KB in flight per CTA, ~7 CTAs per SM: HBM-bound), rows padded by 16 bytes so that the
// thread-per-key score loop reads them without bank conflicts.  The general kernel above read each key row with twelve
// 16-byte loads per THREAD at a 3 KB stride (32 sectors per request, half of each used): 0.23 of the HBM roofline.
constexpr int DA_THREADS = 128;
constexpr int DA_MAX_TK = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}

__global__ void __launch_bounds__(DA_THREADS)
decoder_attention_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k, int64_t ldk,
                         const __nv_bfloat16* __restrict__ v, int64_t ldv, const uint8_t* __restrict__ key_pad,
                         __nv_bfloat16* __restrict__ out, int64_t ldo, int Nq, int Tk, int dh, float scale) {
  extern __shared__ __align__(16) uint8_t da_smem[];
  const int pitch = (dh + 8) * 2;                            // bytes per staged row
  uint8_t* Ks = da_smem;                                     // [Tk][dh + 8] bf16
  uint8_t* Vs = Ks + static_cast<size_t>(Tk) * pitch;
  float* qs = reinterpret_cast<float*>(Vs + static_cast<size_t>(Tk) * pitch);   // [SA_NQ][dh], pre-scaled
  float* sc = qs + SA_NQ * dh;                               // [SA_NQ][Tk]
  const int b = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cpr = dh / 8;                                    // 16-byte chunks per row
  const __nv_bfloat16* kb = k + static_cast<int64_t>(b) * Tk * ldk + h * dh;
  const __nv_bfloat16* vb = v + static_cast<int64_t>(b) * Tk * ldv + h * dh;
  const uint32_t ks_u = smem_u32(Ks), vs_u = smem_u32(Vs);
  for (int i = tid; i < Tk * cpr; i += DA_THREADS) {
    const int r = i / cpr, c = i - r * cpr;
    cp_async16(ks_u + r * pitch + c * 16, kb + static_cast<int64_t>(r) * ldk + c * 8);
  }
  for (int i = tid; i < Tk * cpr; i += DA_THREADS) {
    const int r = i / cpr, c = i - r * cpr;
    cp_async16(vs_u + r * pitch + c * 16, vb + static_cast<int64_t>(r) * ldv + c * 8);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < Nq * dh; i += DA_THREADS) {
    const int qi = i / dh, c = i - qi * dh;
    qs[i] = __bfloat162float(q[(static_cast<int64_t>(b) * Nq + qi) * ldq + h * dh + c]) * scale;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // scores: thread per key, every query
  for (int j = tid; j < Tk; j += DA_THREADS) {
    const bool pad = key_pad != nullptr && key_pad[static_cast<int64_t>(b) * Tk + j] != 0;
    float acc[SA_NQ];
#pragma unroll
    for (int qi = 0; qi < SA_NQ; ++qi) acc[qi] = 0.0f;
    const uint4* kr = reinterpret_cast<const uint4*>(Ks + static_cast<size_t>(j) * pitch);
    for (int c = 0; c < cpr; ++c) {
      const uint4 u = kr[c];
      const float kv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y),
                           bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
      for (int qi = 0; qi < SA_NQ; ++qi) {
        if (qi < Nq) {
          const float4 q0 = *reinterpret_cast<const float4*>(qs + qi * dh + c * 8);
          const float4 q1 = *reinterpret_cast<const float4*>(qs + qi * dh + c * 8 + 4);
          acc[qi] = fmaf(q0.x, kv[0], fmaf(q0.y, kv[1], fmaf(q0.z, kv[2], fmaf(q0.w, kv[3], acc[qi]))));
          acc[qi] = fmaf(q1.x, kv[4], fmaf(q1.y, kv[5], fmaf(q1.z, kv[6], fmaf(q1.w, kv[7], acc[qi]))));
        }
      }
    }
#pragma unroll
    for (int qi = 0; qi < SA_NQ; ++qi)
      if (qi < Nq) sc[qi * Tk + j] = pad ? -INFINITY : acc[qi];
  }
  __syncthreads();
  // softmax per query row (a fully masked row gives NaN, like torch.softmax)
  for (int qi = warp; qi < Nq; qi += DA_THREADS / 32) {
    float m = -INFINITY;
    for (int j = lane; j < Tk; j += 32) m = fmaxf(m, sc[qi * Tk + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.0f;
    for (int j = lane; j < Tk; j += 32) {
      const float e = expf(sc[qi * Tk + j] - m);
      sc[qi * Tk + j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int j = lane; j < Tk; j += 32) sc[qi * Tk + j] *= inv;
  }
  __syncthreads();
  // P.V: thread per pair of columns
  for (int c2 = tid; c2 < dh / 2; c2 += DA_THREADS) {
    float a0[SA_NQ], a1[SA_NQ];
#pragma unroll
    for (int qi = 0; qi < SA_NQ; ++qi) a0[qi] = a1[qi] = 0.0f;
    for (int j = 0; j < Tk; ++j) {
      const uint32_t vv = *reinterpret_cast<const uint32_t*>(Vs + static_cast<size_t>(j) * pitch + c2 * 4);
      const float v0 = bf16_lo(vv), v1 = bf16_hi(vv);
#pragma unroll
      for (int qi = 0; qi < SA_NQ; ++qi) {
        if (qi < Nq) {
          const float pr = sc[qi * Tk + j];
          a0[qi] = fmaf(pr, v0, a0[qi]);
          a1[qi] = fmaf(pr, v1, a1[qi]);
        }
      }
    }
#pragma unroll
    for (int qi = 0; qi < SA_NQ; ++qi)
      if (qi < Nq)
        *reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * Nq + qi) * ldo + h * dh + c2 * 2) = pack_bf16(a0[qi], a1[qi]);
  }
}

// ------------------------------------------------------------------ decoder attention, warp-private form (default)
// The CTA-per-(utterance, head) kernel above spends most of a CTA's life in its score / P.V loops with 64 and 48 of its
// 128 threads busy, so on average few of an SM's bytes are in flight: 0.31 of the HBM roofline (ncu launch list v17:
// 456 MB in 224 us).  Here a WARP owns a stream of (utterance, head) items and a private two-slot ring in shared memory:
// while the tensor cores work on slot s (legacy mma.sync m16n8k16 -- the N_q <= 8 query rows fill half of one M = 16
// tile, which is wasteful and irrelevant: ~100 MMAs per item against a 24 KB tile), the 16-byte cp.async copies of
// the next item's K | V | Q land in slot s ^ 1.  No __syncthreads, no barrier between warps; every lane copies with
// a fixed (row mod R, chunk) pattern so the issue loop is one LDGSTS + pointer bumps per chunk.  Fragments come from
// ldmatrix (K) and ldmatrix.trans (V) over rows padded by 16 bytes (conflict-free); the softmax runs on the accumulator
// layout (a query row lives in the four lanes of a quad), P goes into P.V as bf16 hi + lo parts.
// A fully masked row gives NaN, like torch.softmax.  grid = one CTA per SM, persistent.
template <int DH, int KT>
struct DecAttn {
  static constexpr int TKP = KT * 8;                 // staged key rows (rows >= T_k stay zero)
  static constexpr int PITCH = DH * 2 + 16;          // bytes per staged row
  static constexpr int CPR = DH / 8;                 // 16-byte chunks per row
  static constexpr int RPI = 32 / CPR;               // rows copied per warp iteration
  static constexpr int SLOT = (2 * TKP + 8) * PITCH; // K | V | Q (8 rows)
  static constexpr int WARPS_FIT = (227 * 1024) / (2 * SLOT);
  static constexpr int WARPS = WARPS_FIT >= 4 ? 4 : WARPS_FIT;
  static_assert(DH % 16 == 0 && KT % 2 == 0 && WARPS >= 1, "decoder attention: unsupported tile");
};

__device__ __forceinline__ void da_ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void da_ldsm_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
// rows 8..15 of the A tile are padding: a1 = a3 = 0
__device__ __forceinline__ void da_mma(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

template <int DH, int KT>
__global__ void __launch_bounds__(DecAttn<DH, KT>::WARPS * 32, 1)
decoder_attention_mma_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k, int64_t ldk,
                             const __nv_bfloat16* __restrict__ v, int64_t ldv, const uint8_t* __restrict__ key_pad,
                             __nv_bfloat16* __restrict__ out, int64_t ldo, int B, int H, int Nq, int Tk, float scale) {
  using C = DecAttn<DH, KT>;
  constexpr int PW = (C::TKP + 31) / 32;   // 32-key mask words
  extern __shared__ __align__(16) uint8_t dam_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  uint8_t* ring_p = dam_smem + static_cast<size_t>(warp) * 2 * C::SLOT;
  const uint32_t ring = smem_u32(ring_p);
  // rows the copies never write (keys >= T_k, queries >= N_q) are zero for the kernel's lifetime
  for (int s = 0; s < 2; ++s) {
    for (int i = lane; i < (C::TKP - Tk) * (C::PITCH / 16); i += 32) {
      reinterpret_cast<uint4*>(ring_p + s * C::SLOT + Tk * C::PITCH)[i] = make_uint4(0u, 0u, 0u, 0u);
      reinterpret_cast<uint4*>(ring_p + s * C::SLOT + (C::TKP + Tk) * C::PITCH)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = lane; i < (8 - Nq) * (C::PITCH / 16); i += 32)
      reinterpret_cast<uint4*>(ring_p + s * C::SLOT + (2 * C::TKP + Nq) * C::PITCH)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncwarp();
  const int64_t items = static_cast<int64_t>(B) * H;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * C::WARPS;
  // copy pattern of this lane: row (lane / CPR) of every group of RPI rows, chunk lane % CPR
  const int cr = lane / C::CPR, cc = lane - cr * C::CPR;
  const bool copier = cr < C::RPI;
  uint32_t pad_next[PW];
  auto issue = [&](int64_t item, int slot) {
#pragma unroll
    for (int w = 0; w < PW; ++w) pad_next[w] = 1u;
    if (item < items) {
      const int b = static_cast<int>(item / H), h = static_cast<int>(item - static_cast<int64_t>(b) * H);
      if (copier) {
        const __nv_bfloat16* ks = k + (static_cast<int64_t>(b) * Tk + cr) * ldk + h * DH + cc * 8;
        const __nv_bfloat16* vs = v + (static_cast<int64_t>(b) * Tk + cr) * ldv + h * DH + cc * 8;
        uint32_t dst = ring + slot * C::SLOT + cr * C::PITCH + cc * 16;
        for (int r = cr; r < Tk; r += C::RPI) {
          cp_async16(dst, ks);
          cp_async16(dst + C::TKP * C::PITCH, vs);
          ks += C::RPI * ldk;
          vs += C::RPI * ldv;
          dst += C::RPI * C::PITCH;
        }
        const __nv_bfloat16* qs = q + (static_cast<int64_t>(b) * Nq + cr) * ldq + h * DH + cc * 8;
        dst = ring + slot * C::SLOT + (2 * C::TKP + cr) * C::PITCH + cc * 16;
        for (int r = cr; r < Nq; r += C::RPI) {
          cp_async16(dst, qs);
          qs += C::RPI * ldq;
          dst += C::RPI * C::PITCH;
        }
      }
#pragma unroll
      for (int w = 0; w < PW; ++w) {
        const int j = w * 32 + lane;
        pad_next[w] = j >= Tk ? 1u : (key_pad != nullptr ? static_cast<uint32_t>(__ldg(key_pad + static_cast<int64_t>(b) * Tk + j)) : 0u);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int64_t item = static_cast<int64_t>(blockIdx.x) * C::WARPS + warp;
  int slot = 0;
  issue(item, 0);
  while (item < items) {
    uint32_t padw[PW];
#pragma unroll
    for (int w = 0; w < PW; ++w) padw[w] = __ballot_sync(0xffffffffu, pad_next[w] != 0u);
    __syncwarp();   // every lane is done reading slot ^ 1 (the previous item)
    issue(item + stride, slot ^ 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    const int b = static_cast<int>(item / H), h = static_cast<int>(item - static_cast<int64_t>(b) * H);
    const uint32_t sK = ring + slot * C::SLOT, sV = sK + C::TKP * C::PITCH, sQ = sV + C::TKP * C::PITCH;
    // ---- S = Q K^T: M = queries (rows g), N = keys, K = dh
    float s[KT][4];
#pragma unroll
    for (int nt = 0; nt < KT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
    const int mi = lane >> 3, mr = lane & 7;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
      uint32_t a0, a2;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a0) : "r"(sQ + g * C::PITCH + (ks * 16 + 2 * t4) * 2) : "memory");
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a2) : "r"(sQ + g * C::PITCH + (ks * 16 + 8 + 2 * t4) * 2) : "memory");
#pragma unroll
      for (int nt = 0; nt < KT; nt += 2) {
        uint32_t r0, r1, r2, r3;
        da_ldsm_x4(r0, r1, r2, r3, sK + ((nt + (mi >> 1)) * 8 + mr) * C::PITCH + (ks * 16 + (mi & 1) * 8) * 2);
        da_mma(s[nt], a0, a2, r0, r1);
        da_mma(s[nt + 1], a0, a2, r2, r3);
      }
    }
    // ---- softmax of row g over the keys of this lane's quad (keys nt * 8 + 2 t4 + e)
    float m = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < KT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t4 + e;
        const bool masked = (padw[nt >> 2] >> (j & 31)) & 1u;
        s[nt][e] = masked ? -INFINITY : s[nt][e] * scale;
        m = fmaxf(m, s[nt][e]);
      }
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    float sum = 0.0f;
#pragma unroll
    for (int nt = 0; nt < KT; ++nt) {
      s[nt][0] = expf(s[nt][0] - m);
      s[nt][1] = expf(s[nt][1] - m);
      sum += s[nt][0] + s[nt][1];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    // P as bf16 hi + lo (two MMAs per tile: the tensor core is idle anyway, and the decoder's outputs keep the
    // accuracy of the fp32-probability form this kernel replaces)
    uint32_t pa[KT], pl[KT];
#pragma unroll
    for (int nt = 0; nt < KT; ++nt) {
      const float p0 = s[nt][0] * inv, p1 = s[nt][1] * inv;
      pa[nt] = pack_bf16(p0, p1);
      pl[nt] = pack_bf16(p0 - bf16_lo(pa[nt]), p1 - bf16_hi(pa[nt]));
    }
    // ---- O = P V: M = queries, N = dh, K = keys
    float o[DH / 8][4];
#pragma unroll
    for (int ct = 0; ct < DH / 8; ++ct) o[ct][0] = o[ct][1] = o[ct][2] = o[ct][3] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < KT / 2; ++kk) {
#pragma unroll
      for (int ct = 0; ct < DH / 8; ct += 2) {
        uint32_t r0, r1, r2, r3;
        da_ldsm_x4_trans(r0, r1, r2, r3, sV + (kk * 16 + (mi & 1) * 8 + mr) * C::PITCH + (ct + (mi >> 1)) * 16);
        da_mma(o[ct], pa[2 * kk], pa[2 * kk + 1], r0, r1);
        da_mma(o[ct + 1], pa[2 * kk], pa[2 * kk + 1], r2, r3);
        da_mma(o[ct], pl[2 * kk], pl[2 * kk + 1], r0, r1);
        da_mma(o[ct + 1], pl[2 * kk], pl[2 * kk + 1], r2, r3);
      }
    }
    if (g < Nq) {
      __nv_bfloat16* orow = out + (static_cast<int64_t>(b) * Nq + g) * ldo + h * DH + 2 * t4;
#pragma unroll
      for (int ct = 0; ct < DH / 8; ++ct) *reinterpret_cast<uint32_t*>(orow + ct * 8) = pack_bf16(o[ct][0], o[ct][1]);
    }
    item += stride;
    slot ^= 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ small attention, backward (training)
// Gradients of the decoder attention (models/emotion_decoder.py:42, :48-54; N_q learned queries, T_k keys) for one
// (utterance, head) per CTA.  P is rebuilt from q, k (exact softmax, as in the forward); then
//   dV = P^T dO,   dP = dO V^T,   dS = P * (dP - rowsum(dP * P)),   dQ = scale dS K,   dK = scale dS^T Q.
// Everything is tiny (N_q <= 8, T_k <= a few hundred): fp32 on CUDA cores, shared memory for P / dS.
constexpr int SB_THREADS = 128;

__global__ void __launch_bounds__(SB_THREADS)
small_attention_backward_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k,
                                int64_t ldk, const __nv_bfloat16* __restrict__ v, int64_t ldv,
                                const __nv_bfloat16* __restrict__ dout, int64_t lddo, const uint8_t* __restrict__ key_pad,
                                __nv_bfloat16* __restrict__ dq, int64_t lddq, __nv_bfloat16* __restrict__ dk, int64_t lddk,
                                __nv_bfloat16* __restrict__ dv, int64_t lddv, int H, int Nq, int Tk, int dh, float scale,
                                uint32_t drop_p8, uint32_t drop_key, float drop_scale) {
  extern __shared__ float sm[];
  float* qs = sm;                       // [SA_NQ][dh]  q * scale
  float* dos = qs + SA_NQ * dh;         // [SA_NQ][dh]  dO
  float* ps = dos + SA_NQ * dh;         // [SA_NQ][Tk]  P
  float* dss = ps + SA_NQ * Tk;         // [SA_NQ][Tk]  dP -> dS
  float* pms = dss + SA_NQ * Tk;        // [SA_NQ][Tk]  P o M / (1 - p): what multiplied V in the forward (= P without dropout)
  const int b = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < Nq * dh; i += SB_THREADS) {
    const int qi = i / dh, c = i - qi * dh;
    qs[qi * dh + c] = __bfloat162float(q[(static_cast<int64_t>(b) * Nq + qi) * ldq + h * dh + c]) * scale;
    dos[qi * dh + c] = __bfloat162float(dout[(static_cast<int64_t>(b) * Nq + qi) * lddo + h * dh + c]);
  }
  __syncthreads();
  // scores and dP = dO . v_j, thread per key
  for (int j = tid; j < Tk; j += SB_THREADS) {
    const bool pad = key_pad != nullptr && key_pad[static_cast<int64_t>(b) * Tk + j] != 0;
    const __nv_bfloat16* kr = k + (static_cast<int64_t>(b) * Tk + j) * ldk + h * dh;
    const __nv_bfloat16* vr = v + (static_cast<int64_t>(b) * Tk + j) * ldv + h * dh;
    float s[SA_NQ], dp[SA_NQ];
#pragma unroll
    for (int qi = 0; qi < SA_NQ; ++qi) { s[qi] = 0.0f; dp[qi] = 0.0f; }
    for (int c = 0; c < dh; ++c) {
      const float kv = __bfloat162float(kr[c]), vv = __bfloat162float(vr[c]);
#pragma unroll
      for (int qi = 0; qi < SA_NQ; ++qi) {
        if (qi < Nq) {
          s[qi] = fmaf(qs[qi * dh + c], kv, s[qi]);
          dp[qi] = fmaf(dos[qi * dh + c], vv, dp[qi]);
        }
      }
    }
#pragma unroll
    for (int qi = 0; qi < SA_NQ; ++qi) {
      if (qi < Nq) {
        ps[qi * Tk + j] = pad ? -INFINITY : s[qi];
        dss[qi * Tk + j] = dp[qi];
      }
    }
  }
  __syncthreads();
  // softmax per query row, then dS = P * (dP - sum_j dP_j P_j)  (warp per row)
  for (int qi = warp; qi < Nq; qi += SB_THREADS / 32) {
    float m = -INFINITY;
    for (int j = lane; j < Tk; j += 32) m = fmaxf(m, ps[qi * Tk + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.0f;
    for (int j = lane; j < Tk; j += 32) {
      const float e = expf(ps[qi * Tk + j] - m);
      ps[qi * Tk + j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    float dot = 0.0f;
    const uint32_t dkey = drop_key_bh(drop_key, static_cast<uint32_t>(b * H + h));
    for (int j = lane; j < Tk; j += 32) {
      const float pr = ps[qi * Tk + j] * inv;
      // O = (P o M c) V: dV takes P o M c, dP = (dO V^T) o M c, dS = P o (dP - rowsum(dP o P)) with the undropped P
      const float mc = (drop_p8 == 0u || drop_keep(dkey, static_cast<uint32_t>(qi), static_cast<uint32_t>(j), drop_p8)) ? drop_scale : 0.0f;
      const float dpm = dss[qi * Tk + j] * mc;
      ps[qi * Tk + j] = pr;
      pms[qi * Tk + j] = pr * mc;
      dss[qi * Tk + j] = dpm;
      dot += pr * dpm;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int j = lane; j < Tk; j += 32) dss[qi * Tk + j] = ps[qi * Tk + j] * (dss[qi * Tk + j] - dot);
  }
  __syncthreads();
  // dV_j = sum_q P_qj dO_q ; dK_j = sum_q dS_qj (q_q * scale)   (thread per (key, column))
  for (int i = tid; i < Tk * dh; i += SB_THREADS) {
    const int j = i / dh, c = i - j * dh;
    float av = 0.0f, ak = 0.0f;
#pragma unroll
    for (int qi = 0; qi < SA_NQ; ++qi) {
      if (qi < Nq) {
        av = fmaf(pms[qi * Tk + j], dos[qi * dh + c], av);
        ak = fmaf(dss[qi * Tk + j], qs[qi * dh + c], ak);
      }
    }
    dv[(static_cast<int64_t>(b) * Tk + j) * lddv + h * dh + c] = __float2bfloat16_rn(av);
    dk[(static_cast<int64_t>(b) * Tk + j) * lddk + h * dh + c] = __float2bfloat16_rn(ak);
  }
  // dQ_q = scale * sum_j dS_qj k_j   (thread per (query, column))
  for (int i = tid; i < Nq * dh; i += SB_THREADS) {
    const int qi = i / dh, c = i - qi * dh;
    float a = 0.0f;
    for (int j = 0; j < Tk; ++j)
      a = fmaf(dss[qi * Tk + j], __bfloat162float(k[(static_cast<int64_t>(b) * Tk + j) * ldk + h * dh + c]), a);
    dq[(static_cast<int64_t>(b) * Nq + qi) * lddq + h * dh + c] = __float2bfloat16_rn(a * scale);
  }
}

// The same gradients with the head's K and V tiles staged in shared memory (the decoder's shapes: T_k <= 128).  The kernel
// above reads each key row element by element from global memory, one row per thread (32 sectors per 2-byte warp load), and
// runs dQ on N_q * dh threads over all keys: 362 us per 512-utterance cross-attention call for ~200 MB (0.09 of HBM, ncu
// training launch list v20).  Here: 16-byte cp.async copies of K | V (coalesced, issued back to back), two threads per key
// for S and dP (halves of dh, combined by one shuffle), 16-byte vector reads of the staged rows, dV / dK written as 16-byte
// chunks (thread per (key, chunk)), dQ as thread per (query, column pair) over the staged K.
constexpr int DB_THREADS = 128;

__global__ void __launch_bounds__(DB_THREADS)
decoder_attention_backward_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k,
                                  int64_t ldk, const __nv_bfloat16* __restrict__ v, int64_t ldv,
                                  const __nv_bfloat16* __restrict__ dout, int64_t lddo, const uint8_t* __restrict__ key_pad,
                                  __nv_bfloat16* __restrict__ dq, int64_t lddq, __nv_bfloat16* __restrict__ dk, int64_t lddk,
                                  __nv_bfloat16* __restrict__ dv, int64_t lddv, int H, int Nq, int Tk, int dh, float scale,
                                  uint32_t drop_p8, uint32_t drop_key, float drop_scale) {
  extern __shared__ __align__(16) uint8_t db_smem[];
  const int pitch = (dh + 8) * 2;                            // bytes per staged row
  const int cpr = dh / 8;                                    // 16-byte chunks per row
  uint8_t* Ks = db_smem;                                     // [Tk][dh + 8] bf16
  uint8_t* Vs = Ks + static_cast<size_t>(Tk) * pitch;
  float* qs = reinterpret_cast<float*>(Vs + static_cast<size_t>(Tk) * pitch);   // [SA_NQ][dh]  q * scale
  float* dos = qs + SA_NQ * dh;         // [SA_NQ][dh]  dO
  float* ps = dos + SA_NQ * dh;         // [SA_NQ][Tk]  P
  float* dss = ps + SA_NQ * Tk;         // [SA_NQ][Tk]  dP -> dS
  float* pms = dss + SA_NQ * Tk;        // [SA_NQ][Tk]  P o M / (1 - p)
  const int b = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* kb = k + static_cast<int64_t>(b) * Tk * ldk + h * dh;
  const __nv_bfloat16* vb = v + static_cast<int64_t>(b) * Tk * ldv + h * dh;
  const uint32_t ks_u = smem_u32(Ks), vs_u = smem_u32(Vs);
  for (int i = tid; i < Tk * cpr; i += DB_THREADS) {
    const int r = i / cpr, c = i - r * cpr;
    cp_async16(ks_u + r * pitch + c * 16, kb + static_cast<int64_t>(r) * ldk + c * 8);
    cp_async16(vs_u + r * pitch + c * 16, vb + static_cast<int64_t>(r) * ldv + c * 8);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < SA_NQ * dh; i += DB_THREADS) {
    const int qi = i / dh, c = i - qi * dh;
    const bool on = qi < Nq;
    qs[i] = on ? __bfloat162float(q[(static_cast<int64_t>(b) * Nq + qi) * ldq + h * dh + c]) * scale : 0.0f;
    dos[i] = on ? __bfloat162float(dout[(static_cast<int64_t>(b) * Nq + qi) * lddo + h * dh + c]) : 0.0f;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // scores and dP = dO . v_j: two threads per key (chunks [0, cpr/2) and [cpr/2, cpr)), combined by one shuffle
  const int c_half = cpr / 2;
  for (int j2 = tid; j2 < ((2 * Tk + 31) & ~31); j2 += DB_THREADS) {
    const int j = j2 >> 1, half = j2 & 1;
    float s[SA_NQ], dp[SA_NQ];
#pragma unroll
    for (int qi = 0; qi < SA_NQ; ++qi) { s[qi] = 0.0f; dp[qi] = 0.0f; }
    if (j < Tk) {
      const uint4* kr = reinterpret_cast<const uint4*>(Ks + static_cast<size_t>(j) * pitch);
      const uint4* vr = reinterpret_cast<const uint4*>(Vs + static_cast<size_t>(j) * pitch);
      const int c0 = half ? c_half : 0, c1 = half ? cpr : c_half;
      for (int c = c0; c < c1; ++c) {
        const uint4 ku = kr[c], vu = vr[c];
        const float kv[8] = {bf16_lo(ku.x), bf16_hi(ku.x), bf16_lo(ku.y), bf16_hi(ku.y), bf16_lo(ku.z), bf16_hi(ku.z), bf16_lo(ku.w), bf16_hi(ku.w)};
        const float vv[8] = {bf16_lo(vu.x), bf16_hi(vu.x), bf16_lo(vu.y), bf16_hi(vu.y), bf16_lo(vu.z), bf16_hi(vu.z), bf16_lo(vu.w), bf16_hi(vu.w)};
#pragma unroll
        for (int qi = 0; qi < SA_NQ; ++qi) {
          if (qi < Nq) {
            const float4 q0 = *reinterpret_cast<const float4*>(qs + qi * dh + c * 8), q1 = *reinterpret_cast<const float4*>(qs + qi * dh + c * 8 + 4);
            const float4 d0 = *reinterpret_cast<const float4*>(dos + qi * dh + c * 8), d1 = *reinterpret_cast<const float4*>(dos + qi * dh + c * 8 + 4);
            s[qi] = fmaf(q0.x, kv[0], fmaf(q0.y, kv[1], fmaf(q0.z, kv[2], fmaf(q0.w, kv[3], s[qi]))));
            s[qi] = fmaf(q1.x, kv[4], fmaf(q1.y, kv[5], fmaf(q1.z, kv[6], fmaf(q1.w, kv[7], s[qi]))));
            dp[qi] = fmaf(d0.x, vv[0], fmaf(d0.y, vv[1], fmaf(d0.z, vv[2], fmaf(d0.w, vv[3], dp[qi]))));
            dp[qi] = fmaf(d1.x, vv[4], fmaf(d1.y, vv[5], fmaf(d1.z, vv[6], fmaf(d1.w, vv[7], dp[qi]))));
          }
        }
      }
    }
#pragma unroll
    for (int qi = 0; qi < SA_NQ; ++qi) {
      s[qi] += __shfl_xor_sync(0xffffffffu, s[qi], 1);
      dp[qi] += __shfl_xor_sync(0xffffffffu, dp[qi], 1);
    }
    if (j < Tk && half == 0) {
      const bool pad = key_pad != nullptr && key_pad[static_cast<int64_t>(b) * Tk + j] != 0;
#pragma unroll
      for (int qi = 0; qi < SA_NQ; ++qi) {
        if (qi < Nq) {
          ps[qi * Tk + j] = pad ? -INFINITY : s[qi];
          dss[qi * Tk + j] = dp[qi];
        }
      }
    }
  }
  __syncthreads();
  // softmax per query row, then dS = P * (dP - sum_j dP_j P_j)  (warp per row)
  for (int qi = warp; qi < Nq; qi += DB_THREADS / 32) {
    float m = -INFINITY;
    for (int j = lane; j < Tk; j += 32) m = fmaxf(m, ps[qi * Tk + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.0f;
    for (int j = lane; j < Tk; j += 32) {
      const float e = expf(ps[qi * Tk + j] - m);
      ps[qi * Tk + j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    float dot = 0.0f;
    const uint32_t dkey = drop_key_bh(drop_key, static_cast<uint32_t>(b * H + h));
    for (int j = lane; j < Tk; j += 32) {
      const float pr = ps[qi * Tk + j] * inv;
      const float mc = (drop_p8 == 0u || drop_keep(dkey, static_cast<uint32_t>(qi), static_cast<uint32_t>(j), drop_p8)) ? drop_scale : 0.0f;
      const float dpm = dss[qi * Tk + j] * mc;
      ps[qi * Tk + j] = pr;
      pms[qi * Tk + j] = pr * mc;
      dss[qi * Tk + j] = dpm;
      dot += pr * dpm;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int j = lane; j < Tk; j += 32) dss[qi * Tk + j] = ps[qi * Tk + j] * (dss[qi * Tk + j] - dot);
  }
  __syncthreads();
  // dV_j = sum_q (P o M c)_qj dO_q ; dK_j = sum_q dS_qj (q_q * scale)   (thread per (key, 8-column chunk), 16-byte stores)
  for (int i = tid; i < Tk * cpr; i += DB_THREADS) {
    const int j = i / cpr, c = i - j * cpr;
    float av[8], ak[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) av[e] = ak[e] = 0.0f;
#pragma unroll
    for (int qi = 0; qi < SA_NQ; ++qi) {
      if (qi < Nq) {
        const float pm = pms[qi * Tk + j], ds = dss[qi * Tk + j];
        const float4 d0 = *reinterpret_cast<const float4*>(dos + qi * dh + c * 8), d1 = *reinterpret_cast<const float4*>(dos + qi * dh + c * 8 + 4);
        const float4 q0 = *reinterpret_cast<const float4*>(qs + qi * dh + c * 8), q1 = *reinterpret_cast<const float4*>(qs + qi * dh + c * 8 + 4);
        av[0] = fmaf(pm, d0.x, av[0]); av[1] = fmaf(pm, d0.y, av[1]); av[2] = fmaf(pm, d0.z, av[2]); av[3] = fmaf(pm, d0.w, av[3]);
        av[4] = fmaf(pm, d1.x, av[4]); av[5] = fmaf(pm, d1.y, av[5]); av[6] = fmaf(pm, d1.z, av[6]); av[7] = fmaf(pm, d1.w, av[7]);
        ak[0] = fmaf(ds, q0.x, ak[0]); ak[1] = fmaf(ds, q0.y, ak[1]); ak[2] = fmaf(ds, q0.z, ak[2]); ak[3] = fmaf(ds, q0.w, ak[3]);
        ak[4] = fmaf(ds, q1.x, ak[4]); ak[5] = fmaf(ds, q1.y, ak[5]); ak[6] = fmaf(ds, q1.z, ak[6]); ak[7] = fmaf(ds, q1.w, ak[7]);
      }
    }
    uint4 ov, ok;
    ov.x = pack_bf16(av[0], av[1]); ov.y = pack_bf16(av[2], av[3]); ov.z = pack_bf16(av[4], av[5]); ov.w = pack_bf16(av[6], av[7]);
    ok.x = pack_bf16(ak[0], ak[1]); ok.y = pack_bf16(ak[2], ak[3]); ok.z = pack_bf16(ak[4], ak[5]); ok.w = pack_bf16(ak[6], ak[7]);
    *reinterpret_cast<uint4*>(dv + (static_cast<int64_t>(b) * Tk + j) * lddv + h * dh + c * 8) = ov;
    *reinterpret_cast<uint4*>(dk + (static_cast<int64_t>(b) * Tk + j) * lddk + h * dh + c * 8) = ok;
  }
  // dQ_q = scale * sum_j dS_qj k_j   (thread per (query, column pair) over the staged K)
  for (int i = tid; i < Nq * (dh / 2); i += DB_THREADS) {
    const int qi = i / (dh / 2), c2 = i - qi * (dh / 2);
    float a0 = 0.0f, a1 = 0.0f;
    for (int j = 0; j < Tk; ++j) {
      const uint32_t kk = *reinterpret_cast<const uint32_t*>(Ks + static_cast<size_t>(j) * pitch + c2 * 4);
      const float ds = dss[qi * Tk + j];
      a0 = fmaf(ds, bf16_lo(kk), a0);
      a1 = fmaf(ds, bf16_hi(kk), a1);
    }
    *reinterpret_cast<uint32_t*>(dq + (static_cast<int64_t>(b) * Nq + qi) * lddq + h * dh + c2 * 2) = pack_bf16(a0 * scale, a1 * scale);
  }
}

// ------------------------------------------------------------------ post-path outputs
// probs = sigmoid(logits); decisions = probs >= threshold[class] (per-class calibrated thresholds,
// or 0.5 when none are given: equivalent to logit > 0 up to the tie at 0).
__global__ void emotion_outputs_kernel(const float* __restrict__ logits, const float* __restrict__ thresholds,
                                       float* __restrict__ probs, uint8_t* __restrict__ decisions, int64_t total,
                                       int n_classes) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float pr = 1.0f / (1.0f + expf(-logits[i]));
    if (probs != nullptr) probs[i] = pr;
    if (decisions != nullptr) {
      const float th = thresholds != nullptr ? thresholds[i % n_classes] : 0.5f;
      decisions[i] = pr >= th ? 1 : 0;   // NaN logits -> NaN prob -> 0, as numpy's >= does
    }
  }
}

static int launch_small_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                  int64_t ldv, const uint8_t* key_pad, void* out, int64_t ldo, float* probs,
                                  int B, int H, int Nq, int Tk, int dh, float scale, cudaStream_t s, uint32_t drop_p8 = 0,
                                  uint32_t drop_key = 0, float drop_scale = 1.0f) {
  const bool decoder_shape = drop_p8 == 0u && probs == nullptr && out != nullptr && v != nullptr && Nq <= SA_NQ && Tk <= DA_MAX_TK &&
                             dh <= 128 && ldv % 8 == 0 && ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(v) & 15u) == 0 &&
                             (reinterpret_cast<uintptr_t>(out) & 3u) == 0;
  static const bool force_v1 = getenv("HRIEMO_DECODER_ATTN_V1") != nullptr;   // the CTA-per-(utterance, head) form, for A / B runs
  if (decoder_shape && !force_v1 && (dh == 64 || dh == 96 || dh == 128) && ldq % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(q) & 15u) == 0) {
    // warp-private ring + mma.sync (decoder_attention_mma_kernel)
    using bf = __nv_bfloat16;
    cudaError_t e = cudaSuccess;
    const int64_t items = static_cast<int64_t>(B) * H;
#define HRIEMO_DA_LAUNCH(DHV, KTV)                                                                                          \
    do {                                                                                                                    \
      using C = DecAttn<DHV, KTV>;                                                                                          \
      constexpr int smem_bytes = C::WARPS * 2 * C::SLOT;                                                                    \
      static uint64_t attr_done = 0;                                                                                        \
      if (device_needs_attr(&attr_done))                                                                                    \
        e = cudaFuncSetAttribute(decoder_attention_mma_kernel<DHV, KTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes); \
      if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "small_attention: %s", cudaGetErrorString(e));               \
      int64_t grid = (items + C::WARPS - 1) / C::WARPS;                                                                     \
      if (grid > device_sm_count()) grid = device_sm_count();                                                               \
      decoder_attention_mma_kernel<DHV, KTV><<<static_cast<unsigned>(grid), C::WARPS * 32, smem_bytes, s>>>(                \
          static_cast<const bf*>(q), ldq, static_cast<const bf*>(k), ldk, static_cast<const bf*>(v), ldv, key_pad,          \
          static_cast<bf*>(out), ldo, B, H, Nq, Tk, scale);                                                                 \
    } while (0)
    if (Tk <= 64) {
      if (dh == 64) HRIEMO_DA_LAUNCH(64, 8); else if (dh == 96) HRIEMO_DA_LAUNCH(96, 8); else HRIEMO_DA_LAUNCH(128, 8);
    } else {
      if (dh == 64) HRIEMO_DA_LAUNCH(64, 16); else if (dh == 96) HRIEMO_DA_LAUNCH(96, 16); else HRIEMO_DA_LAUNCH(128, 16);
    }
#undef HRIEMO_DA_LAUNCH
    return check_launch("small_attention");
  }
  if (decoder_shape && H <= 65535) {
    // the decoder's own shapes: K / V staged in shared memory (see decoder_attention_kernel)
    const size_t sm = static_cast<size_t>(2) * Tk * (dh + 8) * 2 + sizeof(float) * (static_cast<size_t>(SA_NQ) * dh + static_cast<size_t>(SA_NQ) * Tk);
    static uint64_t da_attr_done = 0;
    if (sm > 48 * 1024 && device_needs_attr(&da_attr_done)) {
      cudaError_t e = cudaFuncSetAttribute(decoder_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "small_attention: %s", cudaGetErrorString(e));
    }
    dim3 grid(B, H);
    decoder_attention_kernel<<<grid, DA_THREADS, sm, s>>>(
        static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k), ldk,
        static_cast<const __nv_bfloat16*>(v), ldv, key_pad, static_cast<__nv_bfloat16*>(out), ldo, Nq, Tk, dh, scale);
    return check_launch("small_attention");
  }
  const size_t smem = sizeof(float) * (static_cast<size_t>(SA_NQ) * dh + static_cast<size_t>(SA_NQ) * Tk * (probs ? 2 : 1));
  if (smem > 200 * 1024) return set_error(HRIEMO_ERR_INVALID, "small_attention: Tk=%d too long", Tk);
  static uint64_t attr_done = 0;
  if (smem > 48 * 1024 && device_needs_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(small_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "small_attention: %s", cudaGetErrorString(e));
  }
  dim3 grid(B, (Nq + SA_NQ - 1) / SA_NQ, (probs == nullptr && H <= 65535) ? H : 1);
  small_attention_kernel<<<grid, SA_THREADS, smem, s>>>(
      static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k), ldk,
      static_cast<const __nv_bfloat16*>(v), ldv, key_pad, static_cast<__nv_bfloat16*>(out), ldo, probs, H, Nq,
      Tk, dh, scale, drop_p8, drop_key, drop_p8 ? drop_scale : 1.0f);
  return check_launch("small_attention");
}

}  // namespace hriemo

using namespace hriemo;

extern "C" int hriemo_sgemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                                float* out, int64_t ldo, int64_t M, int32_t N, int32_t K, int32_t act,
                                void* stream) {
  HRIEMO_REQUIRE(A && W && out, "sgemm: null pointer");
  HRIEMO_REQUIRE(M >= 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldo >= N, "sgemm: bad shape");
  HRIEMO_REQUIRE(act >= 0 && (act & 3) <= 2 && act <= 6, "sgemm: unknown activation %d", act);
  if (M == 0) return HRIEMO_OK;
  if (N <= SGN_MAX_N && K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15u) == 0 &&
      (reinterpret_cast<uintptr_t>(W) & 15u) == 0) {
    int64_t g = (M + 7) / 8;
    if (g > static_cast<int64_t>(device_sm_count()) * 8) g = static_cast<int64_t>(device_sm_count()) * 8;
    sgemm_few_cols_kernel<<<static_cast<unsigned>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(A, lda, W, ldw, bias, out, ldo, M, N, K, act);
    return check_launch("sgemm_f32 (few columns)");
  }
  HRIEMO_REQUIRE((M + SG_BM - 1) / SG_BM <= 65535, "sgemm: M too large");
  dim3 grid((N + SG_BN - 1) / SG_BN, static_cast<unsigned>((M + SG_BM - 1) / SG_BM));
  sgemm_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, lda, W, ldw, bias, out, ldo, M, N,
                                                                       K, act);
  return check_launch("sgemm_f32");
}

extern "C" int hriemo_small_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                      int64_t ldv, const uint8_t* key_pad, void* out_bf16, int64_t ldo,
                                      float* probs, int32_t B, int32_t H, int32_t Nq, int32_t Tk, int32_t dh,
                                      float scale, void* stream) {
  HRIEMO_REQUIRE(q && k && v && out_bf16, "small_attention: null pointer");
  HRIEMO_REQUIRE(B > 0 && B <= 2147483647 && H > 0 && Nq > 0 && Tk > 0 && dh > 0 && dh % 8 == 0,
                 "small_attention: bad shape");
  HRIEMO_REQUIRE(ldk % 8 == 0 && (reinterpret_cast<uintptr_t>(k) & 15u) == 0, "small_attention: K misaligned");
  return launch_small_attention(q, ldq, k, ldk, v, ldv, key_pad, out_bf16, ldo, probs, B, H, Nq, Tk, dh, scale,
                                static_cast<cudaStream_t>(stream));
}

extern "C" int hriemo_small_attention_dropout(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                              const uint8_t* key_pad, void* out_bf16, int64_t ldo, int32_t B, int32_t H,
                                              int32_t Nq, int32_t Tk, int32_t dh, float scale, uint32_t drop_p8,
                                              uint32_t drop_key, float drop_scale, void* stream) {
  HRIEMO_REQUIRE(q && k && v && out_bf16, "small_attention_dropout: null pointer");
  HRIEMO_REQUIRE(B > 0 && H > 0 && Nq > 0 && Tk > 0 && dh > 0 && dh % 8 == 0, "small_attention_dropout: bad shape");
  HRIEMO_REQUIRE(ldk % 8 == 0 && (reinterpret_cast<uintptr_t>(k) & 15u) == 0, "small_attention_dropout: K misaligned");
  HRIEMO_REQUIRE(drop_p8 <= 255u && (drop_p8 == 0u || drop_scale > 0.0f), "small_attention_dropout: drop_p8 = round(256 p) <= 255");
  return launch_small_attention(q, ldq, k, ldk, v, ldv, key_pad, out_bf16, ldo, nullptr, B, H, Nq, Tk, dh, scale,
                                static_cast<cudaStream_t>(stream), drop_p8, drop_key, drop_scale);
}

extern "C" int hriemo_emotion_outputs(const float* logits, const float* thresholds, float* probs,
                                      uint8_t* decisions, int64_t B, int32_t n_classes, void* stream) {
  HRIEMO_REQUIRE(logits && (probs || decisions) && B >= 0 && n_classes > 0, "emotion_outputs: bad argument");
  if (B == 0) return HRIEMO_OK;
  const int64_t total = B * n_classes;
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  emotion_outputs_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, thresholds, probs, decisions, total, n_classes);
  return check_launch("emotion_outputs");
}

extern "C" int hriemo_attention_probs(const void* q, int64_t ldq, const void* k, int64_t ldk,
                                      const uint8_t* key_pad, float* probs, int32_t B, int32_t H, int32_t Tq,
                                      int32_t Tk, int32_t dh, float scale, void* stream) {
  HRIEMO_REQUIRE(q && k && probs, "attention_probs: null pointer");
  HRIEMO_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0 && dh > 0 && dh % 8 == 0, "attention_probs: bad shape");
  HRIEMO_REQUIRE(ldk % 8 == 0 && (reinterpret_cast<uintptr_t>(k) & 15u) == 0, "attention_probs: K misaligned");
  HRIEMO_REQUIRE((Tq + SA_NQ - 1) / SA_NQ <= 65535, "attention_probs: Tq too large");
  return launch_small_attention(q, ldq, k, ldk, nullptr, 0, key_pad, nullptr, 0, probs, B, H, Tq, Tk, dh, scale,
                                static_cast<cudaStream_t>(stream));
}

static int launch_small_attention_backward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                           const void* d_out, int64_t lddo, const uint8_t* key_pad, void* dq, int64_t lddq,
                                           void* dk, int64_t lddk, void* dv, int64_t lddv, int32_t B, int32_t H, int32_t Nq,
                                           int32_t Tk, int32_t dh, float scale, uint32_t drop_p8, uint32_t drop_key,
                                           float drop_scale, void* stream) {
  HRIEMO_REQUIRE(q && k && v && d_out && dq && dk && dv, "small_attention_backward: null pointer");
  HRIEMO_REQUIRE(B > 0 && H > 0 && H <= 65535 && Nq > 0 && Nq <= SA_NQ && Tk > 0 && dh > 0,
                 "small_attention_backward: bad shape (at most %d queries)", SA_NQ);
  HRIEMO_REQUIRE(drop_p8 <= 255u && (drop_p8 == 0u || drop_scale > 0.0f), "small_attention_backward: drop_p8 = round(256 p) <= 255");
  const size_t smem = sizeof(float) * (2 * static_cast<size_t>(SA_NQ) * dh + 3 * static_cast<size_t>(SA_NQ) * Tk);
  HRIEMO_REQUIRE(smem <= 200 * 1024, "small_attention_backward: Tk=%d too long", Tk);
  static uint64_t attr_done = 0;
  if (smem > 48 * 1024 && device_needs_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(small_attention_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "small_attention_backward: %s", cudaGetErrorString(e));
  }
  using bf = __nv_bfloat16;
  static const bool force_v1 = getenv("HRIEMO_DECODER_ATTN_BWD_V1") != nullptr;   // the unstaged form, for A / B runs
  auto al = [](const void* p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & m) == 0; };
  if (!force_v1 && Tk <= DA_MAX_TK && dh % 8 == 0 && dh <= 128 && ldk % 8 == 0 && ldv % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 &&
      lddq % 2 == 0 && al(k, 15) && al(v, 15) && al(dk, 15) && al(dv, 15) && al(dq, 3)) {
    // K | V staged in shared memory (decoder_attention_backward_kernel)
    const size_t sm2 = static_cast<size_t>(2) * Tk * (dh + 8) * 2 + sizeof(float) * (2 * static_cast<size_t>(SA_NQ) * dh + 3 * static_cast<size_t>(SA_NQ) * Tk);
    static uint64_t db_attr_done = 0;
    if (device_needs_attr(&db_attr_done)) {
      cudaError_t e = cudaFuncSetAttribute(decoder_attention_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "small_attention_backward: %s", cudaGetErrorString(e));
    }
    decoder_attention_backward_kernel<<<dim3(B, H), DB_THREADS, sm2, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf*>(q), ldq, static_cast<const bf*>(k), ldk, static_cast<const bf*>(v), ldv,
        static_cast<const bf*>(d_out), lddo, key_pad, static_cast<bf*>(dq), lddq, static_cast<bf*>(dk), lddk,
        static_cast<bf*>(dv), lddv, H, Nq, Tk, dh, scale, drop_p8, drop_key, drop_p8 ? drop_scale : 1.0f);
    return check_launch("small_attention_backward");
  }
  small_attention_backward_kernel<<<dim3(B, H), SB_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(q), ldq, static_cast<const bf*>(k), ldk, static_cast<const bf*>(v), ldv,
      static_cast<const bf*>(d_out), lddo, key_pad, static_cast<bf*>(dq), lddq, static_cast<bf*>(dk), lddk,
      static_cast<bf*>(dv), lddv, H, Nq, Tk, dh, scale, drop_p8, drop_key, drop_p8 ? drop_scale : 1.0f);
  return check_launch("small_attention_backward");
}

extern "C" int hriemo_small_attention_backward(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                               int64_t ldv, const void* d_out, int64_t lddo, const uint8_t* key_pad,
                                               void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                               int32_t B, int32_t H, int32_t Nq, int32_t Tk, int32_t dh, float scale,
                                               void* stream) {
  return launch_small_attention_backward(q, ldq, k, ldk, v, ldv, d_out, lddo, key_pad, dq, lddq, dk, lddk, dv, lddv, B, H, Nq,
                                         Tk, dh, scale, 0u, 0u, 1.0f, stream);
}

extern "C" int hriemo_small_attention_backward_dropout(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                                       int64_t ldv, const void* d_out, int64_t lddo, const uint8_t* key_pad,
                                                       void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                                       int32_t B, int32_t H, int32_t Nq, int32_t Tk, int32_t dh, float scale,
                                                       uint32_t drop_p8, uint32_t drop_key, float drop_scale, void* stream) {
  return launch_small_attention_backward(q, ldq, k, ldk, v, ldv, d_out, lddo, key_pad, dq, lddq, dk, lddk, dv, lddv, B, H, Nq,
                                         Tk, dh, scale, drop_p8, drop_key, drop_scale, stream);
}
