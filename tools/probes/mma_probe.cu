// Micro-probe: what does the tcgen05 tensor pipe really sustain on the attention kernel's MMA shapes?
// One CTA per SM, one elected thread issues streams of tcgen05.mma (M = 128, K = 16, bf16) and times them with
// clock64(): N = 64 / 96 / 128 / 256, accumulating into ONE tensor-memory tile (a dependent chain, as S = Q K^T over
// dh / 16 steps and O += P V over 4 steps are) or alternating between two, A from shared memory (SS) or tensor memory (TS).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hri-emo_b200/csrc -o tools/probes/mma_probe tools/probes/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace hriemo;

template <int n, int per_group, int groups, int accs, bool ts, bool bmn>
__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 32768, bar = base + 32768 + 65536, slot = bar + 64;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  // fill the operand tiles (values do not matter for timing, NaNs might)
  for (uint32_t i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *slot_ptr;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, n) | (bmn ? kUmmaBMajorMN : 0u);
    const uint64_t a_desc = umma_desc_sw128(sA);
    const uint64_t b_desc = bmn ? umma_desc_mn_sw64(sB, 4096) : umma_desc_sw128(sB);
    uint32_t phase = 0;
    long long total = 0, first = 0;
    for (int rep = 0; rep < 20; ++rep) {
      const long long t0 = clock64();
#pragma unroll
      for (int gq = 0; gq < groups; ++gq) {
        const uint32_t d = tmem + (accs > 1 ? (gq % accs) * 256 : 0);
#pragma unroll
        for (int i = 0; i < per_group; ++i) {
          if (ts) umma_bf16_ts(d, tmem + 448 + (i & 3) * 8, b_desc + (bmn ? ((i & 3) * 16 * 64) >> 4 : 2 * (i & 3)), idesc, i != 0);
          else umma_bf16(d, a_desc + 2 * (i & 3), b_desc + (bmn ? ((i & 3) * 16 * 64) >> 4 : 2 * (i & 3)), idesc, i != 0);
        }
      }
      umma_commit(bar);
      mbar_wait(bar, phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (rep == 0) first = t1 - t0; else total += t1 - t0;
    }
    if (blockIdx.x == 0) { out[0] = total / 19; out[1] = first; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc<512>(tmem); }
}

template <int n, int per_group, int groups, int accs, bool ts, bool bmn>
void run(const char* name, long long* out) {
  cudaFuncSetAttribute(probe<n, per_group, groups, accs, ts, bmn>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  probe<n, per_group, groups, accs, ts, bmn><<<148, 128, 110 * 1024>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  const double mmas = double(per_group) * groups;
  printf("%-66s %8lld cycles = %6.1f per MMA (floor %3.0f)  [%s]\n", name, h[0], h[0] / mmas, 128.0 * n / 256.0, cudaGetErrorString(e));
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  run<64, 6, 1, 1, false, false>("S  N=64  SS  6 dependent MMAs, 1 group (latency)", out);
  run<64, 6, 32, 1, false, false>("S  N=64  SS  6 x 32 groups, one accumulator", out);
  run<64, 6, 32, 2, false, false>("S  N=64  SS  6 x 32 groups, two accumulators alternating", out);
  run<128, 6, 1, 1, false, false>("S  N=128 SS  6 dependent MMAs, 1 group (latency)", out);
  run<128, 6, 32, 1, false, false>("S  N=128 SS  6 x 32 groups, one accumulator", out);
  run<128, 6, 32, 2, false, false>("S  N=128 SS  6 x 32 groups, two accumulators", out);
  run<256, 4, 32, 1, false, false>("S  N=256 SS  4 x 32 groups, one accumulator (GEMM shape, 1 CTA)", out);
  run<96, 4, 1, 1, true, true>("PV N=96  TS  4 dependent MMAs, 1 group (latency), B MN-major", out);
  run<96, 4, 32, 1, true, true>("PV N=96  TS  4 x 32 groups, one accumulator, B MN-major", out);
  run<96, 4, 32, 2, true, true>("PV N=96  TS  4 x 32 groups, two accumulators, B MN-major", out);
  run<96, 8, 32, 1, true, true>("PV N=96  TS  8 x 32 groups (128-key step), one accumulator", out);
  run<96, 4, 32, 1, false, true>("PV N=96  SS  4 x 32 groups, one accumulator, B MN-major", out);
  run<128, 4, 32, 1, true, true>("PV N=128 TS  4 x 32 groups, one accumulator, B MN-major", out);
  run<96, 4, 32, 1, true, false>("PV N=96  TS  4 x 32 groups, one accumulator, B K-major", out);
  return 0;
}
