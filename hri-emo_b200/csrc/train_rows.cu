// Backward row kernels of the gate and the decoder head (BASELINE config 5): everything
// between the loss and the encoder outputs that is not a bf16 Linear / LayerNorm / attention (those are
// gemm_wgrad.cu, elementwise.cu, small_ops.cu).
//
//   models/beta_gate_tacfn.py:79-116   w = sigmoid(MLP([a, t, |a-t|, a*t])), a/t = masked means of LN(stream),
//                                       h = w * LN(a)[:, :L] + (1 - w) * LN(t), beta = mean_d(w)
//   models/emotion_decoder.py:127,153   queries broadcast over the batch; logits = Linear(d, 1)(z)
//
// The gate MLP and the head are fp32 in the forward (beta feeds a bit-sensitive decision), so their backward is
// fp32 too: plain CUDA-core loops, a few GFLOP per step at B = 4096 against the 97 TFLOP of the encoder.
// Every reduction runs in a fixed order (no atomics): the training step is deterministic.
#include <cuda_bf16.h>

#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

// ------------------------------------------------------------------------------------------------ fp32 Linear
// y = x W^T + b with x [M, K], W [N, K].  dX[m, k] = sum_n dY[m, n] W[n, k]: one thread per (m, k), W rows
// read coalesced over k, dY[m, :] is a broadcast.
__global__ void __launch_bounds__(256)
linear_dx_f32_kernel(const float* __restrict__ dY, int64_t lddy, const float* __restrict__ W, int64_t ldw,
                     float* __restrict__ dX, int64_t lddx, int64_t M, int N, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t m = blockIdx.y; m < M; m += gridDim.y) {
    if (k < K) {
      const float* g = dY + m * lddy;
      float acc = 0.0f;
      for (int n = 0; n < N; ++n) acc = fmaf(__ldg(g + n), __ldg(W + static_cast<int64_t>(n) * ldw + k), acc);
      dX[m * lddx + k] = acc;
    }
  }
}

// out[n, k] (+)= sum_m G[m, n] * X[m, k]  (dW = dY^T X); with X == nullptr, X[m, k] = 1 and K = 1 (column sums:
// the bias gradient, or any sum over the leading dimension).  Block = 32 columns k x 8 slices of m; the slices
// are combined through shared memory in a fixed order.
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename TG>
__global__ void __launch_bounds__(256)
outer_sum_kernel(const TG* __restrict__ G, int64_t ldg, const float* __restrict__ X, int64_t ldx,
                 float* __restrict__ out, int64_t ldo, int64_t M, int N, int K, int accumulate) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.y;
  const int k = blockIdx.x * 32 + tx;
  float acc = 0.0f;
  if (k < K) {
    if (X != nullptr) {
      for (int64_t m = ty; m < M; m += 8) acc = fmaf(to_f(G[m * ldg + n]), __ldg(X + m * ldx + k), acc);
    } else {
      for (int64_t m = ty; m < M; m += 8) acc += to_f(G[m * ldg + n]);
    }
  }
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && k < K) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][tx];
    float* o = out + static_cast<int64_t>(n) * ldo + k;
    *o = accumulate ? *o + s : s;
  }
}

// Column sums with the columns along the contiguous axis: out[c] (+)= sum_r x[r, c] (x bf16 or fp32).
// Used for the gradient of the emotion queries (sum over the batch of a [B, N_e * d] view).
template <typename T>
__global__ void __launch_bounds__(256)
colsum_rows_kernel(const T* __restrict__ x, int64_t ldx, float* __restrict__ out, int64_t rows, int cols,
                   int accumulate) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float acc = 0.0f;
  if (c < cols)
    for (int64_t r = ty; r < rows; r += 8) acc += to_f(x[r * ldx + c]);
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][tx];
    out[c] = accumulate ? out[c] + s : s;
  }
}

// dx = dy * act'(y) from the activation's OUTPUT y: act 1 = ReLU (y > 0), 2 = sigmoid (y (1 - y)), 0 = identity.
__global__ void act_backward_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                        float* __restrict__ dx, int64_t n, int act) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float g = dy[i], v = y[i];
    dx[i] = act == 1 ? (v > 0.0f ? g : 0.0f) : (act == 2 ? g * v * (1.0f - v) : g);
  }
}

// ------------------------------------------------------------------------------------------------ gate
// g = [a, t, |a - t|, a * t]  (beta_gate_tacfn.py:87-89)  =>  da = dg0 + sign(a - t) dg2 + t dg3,
// dt = dg1 - sign(a - t) dg2 + a dg3, sign(0) = 0 as torch's abs backward.
__global__ void gate_input_backward_kernel(const float* __restrict__ dg, const float* __restrict__ a,
                                           const float* __restrict__ t, float* __restrict__ da,
                                           float* __restrict__ dt, int64_t B, int d) {
  const int64_t n = B * d;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / d;
    const int c = static_cast<int>(i - b * d);
    const float* g = dg + b * 4 * static_cast<int64_t>(d);
    const float av = a[i], tv = t[i];
    const float diff = av - tv;
    const float sg = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);
    da[i] = g[c] + sg * g[2 * d + c] + tv * g[3 * d + c];
    dt[i] = g[d + c] - sg * g[2 * d + c] + av * g[3 * d + c];
  }
}

// 1 / max(1, #valid positions) per utterance (masked_mean's denominator, beta_gate_tacfn.py:20-24); 1 / T without
// a mask.  One warp per utterance.
__global__ void mask_inv_counts_kernel(const uint8_t* __restrict__ pad, int B, int T, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  int cnt = 0;
  if (pad == nullptr) {
    cnt = lane == 0 ? T : 0;
  } else {
    for (int t = lane; t < T; t += 32) cnt += pad[static_cast<int64_t>(warp) * T + t] ? 0 : 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) out[warp] = 1.0f / static_cast<float>(cnt < 1 ? 1 : cnt);
}

// h = w * na[:, :L] + (1 - w) * nt  =>  dw[b, c] = sum_{l < L} dh[b, l, c] (na[b, l, c] - nt[b, l, c]) + dbeta[b] / d
// (beta = mean_d(w), :95).  One thread per pair of columns, rows read as coalesced bf16x2.
__global__ void __launch_bounds__(128)
gate_dw_kernel(const __nv_bfloat16* __restrict__ dh, int64_t lddh, const __nv_bfloat16* __restrict__ na,
               int64_t ldna, int T_a, const __nv_bfloat16* __restrict__ nt, int64_t ldnt,
               const float* __restrict__ dbeta, float* __restrict__ dw, int L, int d) {
  const int b = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (c >= d) return;
  float a0 = 0.0f, a1 = 0.0f;
  for (int l = 0; l < L; ++l) {
    const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dh + (static_cast<int64_t>(b) * L + l) * lddh + c));
    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(na + (static_cast<int64_t>(b) * T_a + l) * ldna + c));
    const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(nt + (static_cast<int64_t>(b) * L + l) * ldnt + c));
    a0 = fmaf(g.x, x.x - y.x, a0);
    a1 = fmaf(g.y, x.y - y.y, a1);
  }
  const float extra = dbeta != nullptr ? dbeta[b] / static_cast<float>(d) : 0.0f;
  dw[static_cast<int64_t>(b) * d + c] = a0 + extra;
  dw[static_cast<int64_t>(b) * d + c + 1] = a1 + extra;
}

// Gradient w.r.t. one LayerNorm-ed stream n [B, T, d] of the gate:
//   dn[b, l, :] = (l < L ? coef[b, :] * dh[b, l, :] : 0) + (valid[b, l] ? dpool[b, :] * inv_cnt[b] : 0)
// coef = w for the audio stream, 1 - w for the text stream (one_minus).  8 columns per thread.
__global__ void __launch_bounds__(256)
gate_stream_grad_kernel(const __nv_bfloat16* __restrict__ dh, int64_t lddh, int L, const float* __restrict__ w,
                        int one_minus, const float* __restrict__ dpool, const uint8_t* __restrict__ pad,
                        const float* __restrict__ inv_cnt, __nv_bfloat16* __restrict__ dn, int64_t lddn, int64_t B,
                        int T, int d) {
  // warp per row (one 64-bit division per ROW, 16-byte loads of w / dpool; the thread-per-chunk form spent its time
  // in two 64-bit divisions and sixteen scalar loads per 16 bytes written: 0.25 of the HBM roofline)
  const int chunks = d / 8;
  const int lane = threadIdx.x & 31;
  const int64_t rows = B * T;
  for (int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5)) {
    const int64_t b = row / T;
    const int l = static_cast<int>(row - b * T);
    const bool blend = l < L;
    const bool valid = pad == nullptr || pad[b * T + l] == 0;
    const float s = valid ? inv_cnt[b] : 0.0f;
    for (int ci = lane; ci < chunks; ci += 32) {
      const int c = ci * 8;
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.0f;
      if (blend) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(dh + (b * L + l) * lddh + c));
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w + b * d + c)), wb = __ldg(reinterpret_cast<const float4*>(w + b * d + c + 4));
        float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        const float g[8] = {bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y), bf16_lo(raw.z), bf16_hi(raw.z), bf16_lo(raw.w), bf16_hi(raw.w)};
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (one_minus ? 1.0f - wv[k] : wv[k]) * g[k];
      }
      if (valid) {
        const float4 pa = __ldg(reinterpret_cast<const float4*>(dpool + b * d + c)), pb = __ldg(reinterpret_cast<const float4*>(dpool + b * d + c + 4));
        const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(pv[k], s, o[k]);
      }
      uint4 outv;
      outv.x = pack_bf16(o[0], o[1]); outv.y = pack_bf16(o[2], o[3]); outv.z = pack_bf16(o[4], o[5]); outv.w = pack_bf16(o[6], o[7]);
      *reinterpret_cast<uint4*>(dn + row * lddn + c) = outv;
    }
  }
}

static unsigned flat_grid_rows(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace hriemo

using namespace hriemo;

extern "C" int hriemo_linear_backward_f32(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* W,
                                          int64_t ldw, int64_t M, int32_t N, int32_t K, float* dX, int64_t lddx,
                                          float* dW, float* db, int32_t accumulate, void* stream) {
  HRIEMO_REQUIRE(dY && M > 0 && N > 0 && K > 0 && lddy >= N, "linear_backward_f32: bad argument");
  HRIEMO_REQUIRE(N <= 65535, "linear_backward_f32: N=%d too large", N);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dX != nullptr) {
    HRIEMO_REQUIRE(W && ldw >= K && lddx >= K, "linear_backward_f32: dX needs W");
    const unsigned gy = static_cast<unsigned>(M < 65535 ? M : 65535);
    linear_dx_f32_kernel<<<dim3((K + 255) / 256, gy), 256, 0, s>>>(dY, lddy, W, ldw, dX, lddx, M, N, K);
    int rc = check_launch("linear_backward_f32 (dX)");
    if (rc) return rc;
  }
  if (dW != nullptr) {
    HRIEMO_REQUIRE(X && ldx >= K, "linear_backward_f32: dW needs X");
    outer_sum_kernel<float><<<dim3((K + 31) / 32, N), 256, 0, s>>>(dY, lddy, X, ldx, dW, K, M, N, K, accumulate);
    int rc = check_launch("linear_backward_f32 (dW)");
    if (rc) return rc;
  }
  if (db != nullptr) {
    outer_sum_kernel<float><<<dim3(1, N), 256, 0, s>>>(dY, lddy, nullptr, 0, db, 1, M, N, 1, accumulate);
    int rc = check_launch("linear_backward_f32 (db)");
    if (rc) return rc;
  }
  return HRIEMO_OK;
}

extern "C" int hriemo_act_backward_f32(const float* dy, const float* y, float* dx, int64_t n, int32_t act,
                                       void* stream) {
  HRIEMO_REQUIRE(dy && y && dx && n >= 0 && act >= 0 && act <= 2, "act_backward_f32: bad argument");
  if (n == 0) return HRIEMO_OK;
  act_backward_f32_kernel<<<flat_grid_rows(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, y, dx, n, act);
  return check_launch("act_backward_f32");
}

extern "C" int hriemo_sum_rows(const void* x, int32_t x_is_f32, int64_t ldx, float* out, int64_t rows, int32_t cols,
                               int32_t accumulate, void* stream) {
  HRIEMO_REQUIRE(x && out && rows > 0 && cols > 0 && ldx >= cols, "sum_rows: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned g = static_cast<unsigned>((cols + 31) / 32);
  if (x_is_f32)
    colsum_rows_kernel<float><<<g, 256, 0, s>>>(static_cast<const float*>(x), ldx, out, rows, cols, accumulate);
  else
    colsum_rows_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ldx, out, rows, cols,
                                                        accumulate);
  return check_launch("sum_rows");
}

extern "C" int hriemo_gate_input_backward(const float* dg, const float* a_pool, const float* t_pool, float* da_pool,
                                          float* dt_pool, int32_t B, int32_t d, void* stream) {
  HRIEMO_REQUIRE(dg && a_pool && t_pool && da_pool && dt_pool && B > 0 && d > 0, "gate_input_backward: bad argument");
  gate_input_backward_kernel<<<flat_grid_rows(static_cast<int64_t>(B) * d), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dg, a_pool, t_pool, da_pool, dt_pool, B, d);
  return check_launch("gate_input_backward");
}

extern "C" int hriemo_mask_inv_counts(const uint8_t* pad, int32_t B, int32_t T, float* inv_counts, void* stream) {
  HRIEMO_REQUIRE(inv_counts && B > 0 && T > 0, "mask_inv_counts: bad argument");
  mask_inv_counts_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(pad, B, T, inv_counts);
  return check_launch("mask_inv_counts");
}

extern "C" int hriemo_gate_blend_backward_w(const void* dh, int64_t lddh, const void* na, int64_t ldna, int32_t T_a,
                                            const void* nt, int64_t ldnt, const float* dbeta, float* dw, int32_t B,
                                            int32_t L, int32_t d, void* stream) {
  HRIEMO_REQUIRE(dh && na && nt && dw && B > 0 && B <= 65535 && L > 0 && T_a >= L && d > 0 && d % 2 == 0,
                 "gate_blend_backward_w: bad shape (B=%d, L=%d, T_a=%d, d=%d)", B, L, T_a, d);
  HRIEMO_REQUIRE(lddh % 2 == 0 && ldna % 2 == 0 && ldnt % 2 == 0 && al16(dh) && al16(na) && al16(nt),
                 "gate_blend_backward_w: misaligned operand");
  using bf = __nv_bfloat16;
  gate_dw_kernel<<<dim3((d / 2 + 127) / 128, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf*>(dh), lddh, static_cast<const bf*>(na), ldna, T_a, static_cast<const bf*>(nt), ldnt, dbeta, dw,
      L, d);
  return check_launch("gate_blend_backward_w");
}

extern "C" int hriemo_gate_stream_grad(const void* dh, int64_t lddh, int32_t L, const float* w, int32_t one_minus,
                                       const float* dpool, const uint8_t* pad, const float* inv_counts, void* dn,
                                       int64_t lddn, int32_t B, int32_t T, int32_t d, void* stream) {
  HRIEMO_REQUIRE(dh && w && dpool && inv_counts && dn && B > 0 && T > 0 && L > 0 && L <= T && d > 0 && d % 8 == 0,
                 "gate_stream_grad: bad shape (B=%d, T=%d, L=%d, d=%d)", B, T, L, d);
  HRIEMO_REQUIRE(lddh % 8 == 0 && lddn % 8 == 0 && al16(dh) && al16(dn) && al16(w) && al16(dpool), "gate_stream_grad: misaligned operand");
  using bf = __nv_bfloat16;
  gate_stream_grad_kernel<<<flat_grid_rows(static_cast<int64_t>(B) * T * (d / 8)), 256, 0,
                            static_cast<cudaStream_t>(stream)>>>(static_cast<const bf*>(dh), lddh, L, w, one_minus, dpool,
                                                                 pad, inv_counts, static_cast<bf*>(dn), lddn, B, T, d);
  return check_launch("gate_stream_grad");
}
