// Loss and optimizer kernels of the training step (BASELINE config 5): the parts of
// scripts/fusion/train_fusion_seq_level_decoder.py:318-335 that are not the model's backward pass.
//
//   loss = BCEWithLogitsLoss()(logits, labels) - 0.01 * mean(beta * (1 - beta))        :319, :326-327
//   clip_grad_norm_(model.parameters(), 5.0)                                             :333
//   AdamW(lr, weight_decay).step()                                                        :334, :405-409
//
// Layout: the 119 parameter tensors of the model (54.6 M elements) are views into ONE flat fp32 arena, and so
// are their gradients and the two AdamW moments.  The global gradient norm is then one reduction, the update one
// launch, and the data-parallel exchange one all-reduce of one buffer.  Nothing here
// synchronises the host: the clip coefficient is computed on the device from the squared norm and read by the
// update kernel from device memory.
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

// One CTA.  loss_out[0] = mean_{b,c} [max(x,0) - x*y + log1p(exp(-|x|))] - beta_weight * mean_b beta(1-beta);
// d_logits = (sigmoid(x) - y) / (B*C);  d_beta = -beta_weight * (1 - 2 beta) / B.   Fixed-order reduction.
__global__ void __launch_bounds__(256)
bce_beta_loss_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                     const float* __restrict__ beta, float beta_weight, int64_t B, int C,
                     float* __restrict__ loss_out, float* __restrict__ d_logits, float* __restrict__ d_beta) {
  __shared__ double part[8];
  const int64_t n = B * C;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = logits[i], y = labels[i];
    acc += static_cast<double>(fmaxf(x, 0.0f) - x * y + log1pf(expf(-fabsf(x))));
    if (d_logits) d_logits[i] = (1.0f / (1.0f + expf(-x)) - y) / static_cast<float>(n);
  }
  acc /= static_cast<double>(n);
  double reg = 0.0;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    const float v = beta[b];
    reg += static_cast<double>(v * (1.0f - v));
    if (d_beta) d_beta[b] = -beta_weight * (1.0f - 2.0f * v) / static_cast<float>(B);
  }
  acc -= static_cast<double>(beta_weight) * reg / static_cast<double>(B);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += part[w];
    loss_out[0] = static_cast<float>(s);
  }
}

// Squared L2 norm of a flat buffer: per-CTA partials in double, fixed order inside a CTA; the final sum over
// the (<= 1184) partials is done by the finalize kernel in a fixed order as well => deterministic.
__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ partials) {
  __shared__ double part[8];
  double acc = 0.0;
  const int64_t n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    acc += static_cast<double>(v.x * v.x + v.y * v.y) + static_cast<double>(v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) acc += static_cast<double>(g[i]) * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += part[w];
    partials[blockIdx.x] = s;
  }
}

// out[0] = total L2 norm, out[1] = clip coefficient min(1, max_norm / (total + 1e-6))  (clip_grad_norm_)
__global__ void clip_finalize_kernel(const double* __restrict__ partials, int n_partials, float max_norm,
                                     float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0;
  for (int i = 0; i < n_partials; ++i) s += partials[i];
  const float total = static_cast<float>(sqrt(s));
  out[0] = total;
  out[1] = max_norm > 0.0f ? fminf(1.0f, max_norm / (total + 1e-6f)) : 1.0f;
}

// torch.optim.AdamW, step t (1-based), over flat fp32 arenas; grad_scale (device, optional) multiplies the
// gradient first (the clip coefficient); w16 (optional) receives the updated parameters as bf16 (the GEMM
// operand copy), so no separate cast pass follows the step.
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n, float decay, float b1, float omb1, float b2, float omb2, float eps, float step_size,
             float bc2_sqrt, const float* __restrict__ grad_scale, __nv_bfloat16* __restrict__ w16) {
  // decay = 1 - lr*wd, omb = 1 - beta, step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t): all formed
  // in double on the host (1 - 0.999f in float is off by 1.3e-5 relative, which lands in exp_avg_sq)
  const float gs = grad_scale != nullptr ? __ldg(grad_scale) : 1.0f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * decay;
    const float mi = b1 * m[i] + omb1 * gi;
    const float vi = b2 * v[i] + omb2 * gi * gi;
    pi -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (w16 != nullptr) w16[i] = __float2bfloat16_rn(pi);
  }
}

static unsigned flat_grid(int64_t n, int per_thread) {
  const int64_t blocks = (n / per_thread + 255) / 256;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;
  return static_cast<unsigned>(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace hriemo

using namespace hriemo;

extern "C" int hriemo_bce_beta_loss(const float* logits, const float* labels, const float* beta, float beta_weight,
                                    int64_t B, int32_t C, float* loss_out, float* d_logits, float* d_beta,
                                    void* stream) {
  HRIEMO_REQUIRE(logits && labels && beta && loss_out && B > 0 && C > 0, "bce_beta_loss: bad argument");
  bce_beta_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, labels, beta, beta_weight, B, C,
                                                                        loss_out, d_logits, d_beta);
  return check_launch("bce_beta_loss");
}

extern "C" int64_t hriemo_grad_norm_workspace_bytes(void) { return static_cast<int64_t>(148 * 8 + 64) * sizeof(double); }

extern "C" int hriemo_grad_norm_clip(const float* grads, int64_t n, float max_norm, void* workspace, float* out2,
                                     void* stream) {
  HRIEMO_REQUIRE(grads && workspace && out2 && n > 0, "grad_norm_clip: bad argument");
  HRIEMO_REQUIRE((reinterpret_cast<uintptr_t>(grads) & 15u) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0,
                 "grad_norm_clip: gradients must be 16-byte aligned, workspace 8-byte aligned");
  unsigned grid = flat_grid(n, 16);
  if (grid > 148u * 8u) grid = 148u * 8u;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* partials = static_cast<double*>(workspace);
  sumsq_partial_kernel<<<grid, 256, 0, s>>>(grads, n, partials);
  int rc = check_launch("grad_norm_clip (partials)");
  if (rc) return rc;
  clip_finalize_kernel<<<1, 32, 0, s>>>(partials, static_cast<int>(grid), max_norm, out2);
  return check_launch("grad_norm_clip (finalize)");
}

extern "C" int hriemo_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 int32_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                                 const float* grad_scale, void* params_bf16, void* stream) {
  HRIEMO_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adamw_step: bad argument");
  HRIEMO_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps > 0.0, "adamw_step: bad hyper-parameter");
  const double bc1 = 1.0 - pow(beta1, step);
  const double bc2s = sqrt(1.0 - pow(beta2, step));
  adamw_kernel<<<flat_grid(n, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, n, static_cast<float>(1.0 - lr * weight_decay), static_cast<float>(beta1),
      static_cast<float>(1.0 - beta1), static_cast<float>(beta2), static_cast<float>(1.0 - beta2),
      static_cast<float>(eps), static_cast<float>(lr / bc1), static_cast<float>(bc2s), grad_scale,
      static_cast<__nv_bfloat16*>(params_bf16));
  return check_launch("adamw_step");
}
