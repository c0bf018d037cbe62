// Kernels of the "tf32-class" precision mode (north_star: logits within 1e-4 of the reference's fp32 forward).
//
// The tensor cores only multiply bf16 here, so an fp32 operand is carried as TWO bf16 numbers, x = hi + lo with
// hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits together; TF32 has 10), and a product as three bf16 products,
//   x . w  ~=  x_hi w_hi + x_lo w_hi + x_hi w_lo            (the dropped lo . lo term is ~2^-16 relative).
// Laid out along K -- activations as [hi | lo | hi], weights as [hi | hi | lo] -- the three terms are ONE run of the
// existing tcgen05 GEMM (gemm_bf16.cu) at 3 K with fp32 accumulation and an fp32 epilogue: no new GEMM kernel.
// What this file adds are the HBM-bound passes around it, all fp32 in / fp32 out:
//   hriemo_split3            fp32 [M, K] -> bf16 [M, 3 K] (activation or weight pattern, optional ReLU on the way)
//   hriemo_attention_f32     softmax(q k^T * scale + key padding) v on CUDA cores, for the encoder's and the decoder's
//                            attention (nn.MultiheadAttention's scaled_dot_product_attention)
//   hriemo_masked_mean_f32   models/beta_gate_tacfn.py:6-24
//   hriemo_gate_blend_f32    models/beta_gate_tacfn.py:95-116
// The mode trades throughput for digits; it is not the benchmark path.
#include <math.h>

#include <cuda_bf16.h>

#include "host_common.h"

namespace hriemo {

// ------------------------------------------------------------------ split3
// pattern 0 (activations): out[r, 0:K) = hi, [K:2K) = lo, [2K:3K) = hi;  pattern 1 (weights): hi, hi, lo
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ldo, int64_t rows,
              int K, int Kp, int pattern, int relu) {
  // Kp = K rounded up to a multiple of 8: each of the three column blocks is Kp wide (zero padded)
  const int64_t total = rows * (Kp / 2);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / (Kp / 2);
    const int c = static_cast<int>(i - r * (Kp / 2)) * 2;
    float v0 = c < K ? x[r * ldx + c] : 0.0f, v1 = c + 1 < K ? x[r * ldx + c + 1] : 0.0f;
    if (relu) { v0 = v0 < 0.0f ? 0.0f : v0; v1 = v1 < 0.0f ? 0.0f : v1; }   // keeps NaN, as torch.relu does
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
    const __nv_bfloat162 hi = __halves2bfloat162(h0, h1), lo = __halves2bfloat162(l0, l1);
    __nv_bfloat16* o = out + r * ldo + c;
    *reinterpret_cast<__nv_bfloat162*>(o) = hi;
    *reinterpret_cast<__nv_bfloat162*>(o + Kp) = pattern == 0 ? lo : hi;
    *reinterpret_cast<__nv_bfloat162*>(o + 2 * Kp) = pattern == 0 ? hi : lo;
  }
}

// ------------------------------------------------------------------ attention, fp32 on CUDA cores
// One CTA per (utterance, head, block of AF_NQ queries): the queries (pre-scaled) and the block's scores live in shared
// memory; keys / values stream from global memory (they are shared by the CTAs of a head and sit in L2).
constexpr int AF_NQ = 16;
constexpr int AF_THREADS = 256;

__global__ void __launch_bounds__(AF_THREADS)
attention_f32_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                     const float* __restrict__ v, int64_t ldv, const uint8_t* __restrict__ key_pad,
                     float* __restrict__ out, int64_t ldo, float* __restrict__ probs, int H, int Tq, int Tk, int dh,
                     float scale) {
  extern __shared__ float sm[];
  float* qs = sm;                    // [AF_NQ][dh]
  float* sc = qs + AF_NQ * dh;       // [AF_NQ][Tkp], Tkp = Tk rounded up to 4 (16-byte rows)
  const int Tkp = (Tk + 3) & ~3;
  const int b = blockIdx.x, h = blockIdx.z;
  const int q0 = blockIdx.y * AF_NQ;
  const int nq = min(AF_NQ, Tq - q0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < nq * dh; i += AF_THREADS) {
    const int qi = i / dh, c = i - qi * dh;
    qs[i] = q[(static_cast<int64_t>(b) * Tq + q0 + qi) * ldq + h * dh + c] * scale;
  }
  __syncthreads();
  // scores: thread per key, all queries of the block
  for (int j = tid; j < Tk; j += AF_THREADS) {
    const bool pad = key_pad != nullptr && key_pad[static_cast<int64_t>(b) * Tk + j] != 0;
    float acc[AF_NQ];
#pragma unroll
    for (int qi = 0; qi < AF_NQ; ++qi) acc[qi] = 0.0f;
    const float4* kr = reinterpret_cast<const float4*>(k + (static_cast<int64_t>(b) * Tk + j) * ldk + h * dh);
    for (int c4 = 0; c4 < dh / 4; ++c4) {
      const float4 kv = __ldg(kr + c4);
#pragma unroll
      for (int qi = 0; qi < AF_NQ; ++qi) {
        const float4 qq = *reinterpret_cast<const float4*>(qs + qi * dh + c4 * 4);
        acc[qi] = fmaf(qq.w, kv.w, fmaf(qq.z, kv.z, fmaf(qq.y, kv.y, fmaf(qq.x, kv.x, acc[qi]))));
      }
    }
#pragma unroll
    for (int qi = 0; qi < AF_NQ; ++qi)
      if (qi < nq) sc[qi * Tkp + j] = pad ? -INFINITY : acc[qi];
  }
  __syncthreads();
  // softmax per query row (a fully masked row gives NaN, like torch.softmax)
  for (int qi = warp; qi < nq; qi += AF_THREADS / 32) {
    float m = -INFINITY;
    for (int j = lane; j < Tk; j += 32) m = fmaxf(m, sc[qi * Tkp + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.0f;
    for (int j = lane; j < Tk; j += 32) {
      const float e = expf(sc[qi * Tkp + j] - m);
      sc[qi * Tkp + j] = e;
      s += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.0f / s;
    for (int j = lane; j < Tk; j += 32) sc[qi * Tkp + j] *= inv;
  }
  __syncthreads();
  if (probs != nullptr) {   // head-averaged probabilities [B, Tq, Tk]: one atomic add per (query, key) and head
    const float inv_h = 1.0f / static_cast<float>(H);
    for (int i = tid; i < nq * Tk; i += AF_THREADS) {
      const int qi = i / Tk, j = i - qi * Tk;
      atomicAdd(probs + (static_cast<int64_t>(b) * Tq + q0 + qi) * Tk + j, sc[qi * Tkp + j] * inv_h);
    }
  }
  // O = P V: thread per (column, group of four queries); four keys per trip (one 16-byte read of each query's
  // probabilities -- a broadcast inside the warp -- and four coalesced reads of V feed sixteen FMAs)
  for (int item = tid; item < dh * (AF_NQ / 4); item += AF_THREADS) {
    const int g = item / dh, c = item - g * dh;
    const float* vr = v + static_cast<int64_t>(b) * Tk * ldv + h * dh + c;
    const float* pr = sc + (g * 4) * Tkp;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int j = 0;
    for (; j + 3 < Tk; j += 4) {
      const float v0 = __ldg(vr + static_cast<int64_t>(j) * ldv), v1 = __ldg(vr + static_cast<int64_t>(j + 1) * ldv);
      const float v2 = __ldg(vr + static_cast<int64_t>(j + 2) * ldv), v3 = __ldg(vr + static_cast<int64_t>(j + 3) * ldv);
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) {
        const float4 p4 = *reinterpret_cast<const float4*>(pr + qi * Tkp + j);
        acc[qi] = fmaf(p4.w, v3, fmaf(p4.z, v2, fmaf(p4.y, v1, fmaf(p4.x, v0, acc[qi]))));
      }
    }
    for (; j < Tk; ++j) {
      const float vv = __ldg(vr + static_cast<int64_t>(j) * ldv);
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) acc[qi] = fmaf(pr[qi * Tkp + j], vv, acc[qi]);
    }
#pragma unroll
    for (int qi = 0; qi < 4; ++qi)
      if (g * 4 + qi < nq) out[(static_cast<int64_t>(b) * Tq + q0 + g * 4 + qi) * ldo + h * dh + c] = acc[qi];
  }
}

// ------------------------------------------------------------------ masked mean over time, fp32
// pooled[b, c] = sum_t x[b, t, c] * valid[b, t] / max(count, 1): CTA per utterance, thread per column, fixed order
__global__ void __launch_bounds__(256)
masked_mean_f32_kernel(const float* __restrict__ x, const uint8_t* __restrict__ pad, float* __restrict__ pooled, int T, int d) {
  const int b = blockIdx.x;
  int count = 0;
  for (int t = 0; t < T; ++t) count += (pad == nullptr || pad[static_cast<int64_t>(b) * T + t] == 0) ? 1 : 0;
  const float inv = 1.0f / static_cast<float>(max(count, 1));
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.0f;
    for (int t = 0; t < T; ++t)
      if (pad == nullptr || pad[static_cast<int64_t>(b) * T + t] == 0) acc += x[(static_cast<int64_t>(b) * T + t) * d + c];
    pooled[static_cast<int64_t>(b) * d + c] = acc * inv;
  }
}

// ------------------------------------------------------------------ gate blend, fp32
// h[b, t, c] = w[b, c] * a[b, t, c] + (1 - w[b, c]) * t_[b, t, c] for t < L (a holds T_a >= L rows per utterance);
// beta[b] = mean_c w[b, c]
__global__ void __launch_bounds__(256)
gate_blend_f32_kernel(const float* __restrict__ a, int T_a, const float* __restrict__ tx, const float* __restrict__ w,
                      float* __restrict__ h, float* __restrict__ beta, int L, int d) {
  const int b = blockIdx.x;
  __shared__ float red[256];
  float s = 0.0f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) s += w[static_cast<int64_t>(b) * d + c];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) beta[b] = red[0] / static_cast<float>(d);
  for (int64_t i = threadIdx.x; i < static_cast<int64_t>(L) * d; i += blockDim.x) {
    const int t = static_cast<int>(i / d), c = static_cast<int>(i - static_cast<int64_t>(t) * d);
    const float wc = w[static_cast<int64_t>(b) * d + c];
    h[(static_cast<int64_t>(b) * L + t) * d + c] =
        wc * a[(static_cast<int64_t>(b) * T_a + t) * d + c] + (1.0f - wc) * tx[(static_cast<int64_t>(b) * L + t) * d + c];
  }
}

}  // namespace hriemo

using namespace hriemo;

extern "C" int hriemo_split3(const float* x, int64_t ldx, void* out_bf16, int64_t ldo, int64_t rows, int32_t K,
                             int32_t pattern, int32_t relu, void* stream) {
  HRIEMO_REQUIRE(x && out_bf16 && rows >= 0 && K > 0 && ldx >= K, "split3: bad argument");
  HRIEMO_REQUIRE(pattern == 0 || pattern == 1, "split3: pattern %d (0 activations [hi|lo|hi], 1 weights [hi|hi|lo])", pattern);
  const int Kp = (K + 7) / 8 * 8;
  HRIEMO_REQUIRE(ldo >= 3 * Kp && ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 3u) == 0,
                 "split3: output needs 3 * roundup(K, 8) columns per row");
  if (rows == 0) return HRIEMO_OK;
  const int64_t total = rows * (Kp / 2);
  int64_t grid = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  if (grid > cap) grid = cap;
  split3_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, static_cast<__nv_bfloat16*>(out_bf16), ldo, rows, K, Kp, pattern, relu);
  return check_launch("split3");
}

extern "C" int hriemo_attention_f32(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                    const uint8_t* key_pad, float* out, int64_t ldo, float* probs, int32_t B, int32_t H,
                                    int32_t Tq, int32_t Tk, int32_t dh, float scale, void* stream) {
  HRIEMO_REQUIRE(q && k && v && out, "attention_f32: null pointer");
  HRIEMO_REQUIRE(B > 0 && H > 0 && H <= 65535 && Tq > 0 && Tk > 0 && dh > 0 && dh % 4 == 0, "attention_f32: bad shape");
  HRIEMO_REQUIRE(ldk % 4 == 0 && (reinterpret_cast<uintptr_t>(k) & 15u) == 0, "attention_f32: K rows must be 16-byte aligned");
  HRIEMO_REQUIRE((Tq + AF_NQ - 1) / AF_NQ <= 65535, "attention_f32: Tq too long");
  const size_t smem = sizeof(float) * (static_cast<size_t>(AF_NQ) * dh + static_cast<size_t>(AF_NQ) * ((Tk + 3) & ~3));
  HRIEMO_REQUIRE(smem <= 200 * 1024, "attention_f32: Tk=%d too long", Tk);
  static uint64_t attr_done = 0;
  if (smem > 48 * 1024 && device_needs_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "attention_f32: %s", cudaGetErrorString(e));
  }
  dim3 grid(B, (Tq + AF_NQ - 1) / AF_NQ, H);
  attention_f32_kernel<<<grid, AF_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(q, ldq, k, ldk, v, ldv, key_pad, out,
                                                                                      ldo, probs, H, Tq, Tk, dh, scale);
  return check_launch("attention_f32");
}

extern "C" int hriemo_masked_mean_f32(const float* x, const uint8_t* pad, float* pooled, int32_t B, int32_t T, int32_t d,
                                      void* stream) {
  HRIEMO_REQUIRE(x && pooled && B > 0 && T > 0 && d > 0, "masked_mean_f32: bad argument");
  masked_mean_f32_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, pad, pooled, T, d);
  return check_launch("masked_mean_f32");
}

extern "C" int hriemo_gate_blend_f32(const float* a, int32_t T_a, const float* t, const float* w, float* h, float* beta,
                                     int32_t B, int32_t L, int32_t d, void* stream) {
  HRIEMO_REQUIRE(a && t && w && h && beta && B > 0 && L > 0 && d > 0 && T_a >= L, "gate_blend_f32: bad argument");
  gate_blend_f32_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, T_a, t, w, h, beta, L, d);
  return check_launch("gate_blend_f32");
}
