#include "host_common.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace hriemo {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return set_error(HRIEMO_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  count_launch();
  return HRIEMO_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(HRIEMO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || (pitch_elems * 2) % 16)
    return set_error(HRIEMO_ERR_INVALID, "TMA operand needs 16-byte aligned base and row pitch");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(HRIEMO_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu pitch=%llu box=%ux%u",
                     (int)r, (unsigned long long)inner, (unsigned long long)outer,
                     (unsigned long long)pitch_elems, box_inner, box_outer);
  return HRIEMO_OK;
}

int make_tmap_bf16_3d_plain(CUtensorMap* map, const void* base, const uint64_t dims[3],
                            const uint64_t pitch_elems[2], const uint32_t box[3], int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(HRIEMO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || (pitch_elems[0] * 2) % 16 || (pitch_elems[1] * 2) % 16)
    return set_error(HRIEMO_ERR_INVALID, "TMA operand needs 16-byte aligned base and pitches");
  cuuint64_t d[3] = {dims[0], dims[1], dims[2]};
  cuuint64_t strides[2] = {pitch_elems[0] * 2, pitch_elems[1] * 2};
  cuuint32_t b[3] = {box[0], box[1], box[2]};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), d, strides, b, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                      : (swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(HRIEMO_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
  return HRIEMO_OK;
}

bool device_needs_attr(uint64_t* done) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (*done & bit) return false;
  *done |= bit;   // benign race: setting the attribute twice is harmless
  return true;
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace hriemo

extern "C" {
int hriemo_version(void) { return HRIEMO_VERSION; }
const char* hriemo_last_error(void) { return hriemo::g_err; }
int64_t hriemo_launch_count(void) { return hriemo::g_launches.load(); }
}
