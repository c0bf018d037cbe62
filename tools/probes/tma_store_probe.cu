// Probe: which inner-dimension start coordinates does a 3-D SWIZZLE_NONE TMA store accept?
// usage: tma_store_probe <t0>   (box 32 x 64 x 1 bf16 into a [B=3][C=64][T=50 (pitch 56)] tensor)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int t0, int b) {
  __shared__ __align__(1024) __nv_bfloat16 box[64 * 32];
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) box[i] = __float2bfloat16(float(i % 32 + 1));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t src = (uint32_t)__cvta_generic_to_shared(box);
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)&tm),
                 "r"(src), "r"(t0), "r"(0), "r"(b) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main(int argc, char** argv) {
  int t0 = atoi(argv[1]);
  const int B = 3, Cc = 64, T = 50, TP = 56;
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  __nv_bfloat16* d; cudaMalloc(&d, B * Cc * TP * 2); cudaMemset(d, 0, B * Cc * TP * 2);
  CUtensorMap tm;
  cuuint64_t dims[3] = {T, Cc, B}; cuuint64_t strides[2] = {TP * 2, (cuuint64_t)Cc * TP * 2};
  cuuint32_t box[3] = {32, 64, 1}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)fnp)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode=%d\n", (int)r);
  k<<<1, 128>>>(tm, t0, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("t0=%d sync: %s\n", t0, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<__nv_bfloat16> h(B * Cc * TP); cudaMemcpy(h.data(), d, h.size() * 2, cudaMemcpyDeviceToHost);
  for (int b = 0; b < B; ++b) { printf("b=%d c=0:", b); for (int t = 0; t < TP; ++t) printf(" %g", __bfloat162float(h[(b * Cc + 0) * TP + t])); printf("\n"); }
  return 0;
}
