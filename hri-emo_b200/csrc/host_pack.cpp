// HOST-side staging helper (no device work): gathers utterances out of the padded fp32 batch the
// reference's collate produced (scripts/fusion/train_fusion_seq_level_decoder.py:191-232) and writes
// them as bf16, trimmed to the slab's own length, into (pinned) staging memory -- half the bytes cross
// PCIe, and none of the padding beyond the slab maximum.  Plain C++ threads; round-to-nearest-even
// exactly like the GPU cast (cvt.rn.bf16x2.f32), so a host-packed slab is bit-identical to a
// device-cast one for every finite input.
#include <stdint.h>
#include <string.h>
#if defined(__GNUC__) && defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/hriemo.h"

namespace hriemo {
int set_error(int code, const char* fmt, ...);   // host_common.cu
}

namespace {

// dst[0..n) = bf16(src[0..n)), round to nearest even, NaN -> 0x7fff (what cvt.rn.bf16.f32 returns);
// written so that gcc vectorises the loop
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx2", "default")))
#endif
void cast_row_scalar(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t u = s[i];
    const uint32_t r = (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
    dst[i] = ((u & 0x7fffffffu) > 0x7f800000u) ? static_cast<uint16_t>(0x7fffu) : static_cast<uint16_t>(r);
  }
}

#if defined(__GNUC__) && defined(__x86_64__)
#define HRIEMO_HAVE_AVX512 1
// The same integer formula on 16 lanes (NOT vcvtneps2bf16: that instruction flushes denormals, the GPU cast does
// not), with NON-TEMPORAL 32-byte stores when the destination is aligned: the pinned staging buffer is written
// once and next read by the DMA engine, so the read-for-ownership of a cached store (2 of 8 bytes of memory
// traffic per element) is pure waste.
__attribute__((target("avx512f,avx512bw,avx512vl")))
void cast_row_avx512(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n) {
  const __m512i c7fff = _mm512_set1_epi32(0x7fff), one = _mm512_set1_epi32(1);
  const __m512i absmask = _mm512_set1_epi32(0x7fffffff), inf = _mm512_set1_epi32(0x7f800000);
  int64_t i = 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31u) == 0;
  for (; i + 16 <= n; i += 16) {
    const __m512i u = _mm512_loadu_si512(reinterpret_cast<const void*>(src + i));
    const __m512i lsb = _mm512_and_si512(_mm512_srli_epi32(u, 16), one);
    __m512i r = _mm512_srli_epi32(_mm512_add_epi32(_mm512_add_epi32(u, c7fff), lsb), 16);
    const __mmask16 nan = _mm512_cmpgt_epu32_mask(_mm512_and_si512(u, absmask), inf);
    r = _mm512_mask_mov_epi32(r, nan, c7fff);
    const __m256i h = _mm512_cvtepi32_epi16(r);
    if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), h);
    else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), h);
  }
  if (i < n) cast_row_scalar(src + i, dst + i, n - i);
}
const bool g_avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                      __builtin_cpu_supports("avx512vl");
#endif

inline void cast_row(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n) {
#ifdef HRIEMO_HAVE_AVX512
  if (g_avx512) return cast_row_avx512(src, dst, n);
#endif
  cast_row_scalar(src, dst, n);
}

struct PackJob {
  const float* src; int64_t ld_src; int64_t T_in; int64_t cols;
  const int32_t* utt; const int32_t* lens;
  uint16_t* dst; int64_t ld_dst; int64_t T_out; int64_t n;
};

// rows [r0, r1) of the [n * T_out] output rows
void pack_rows(const PackJob& j, int64_t r0, int64_t r1) {
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t i = r / j.T_out, t = r - i * j.T_out;
    uint16_t* d = j.dst + r * j.ld_dst;
    const int64_t valid = j.lens ? std::min<int64_t>(j.lens[i], j.T_in) : j.T_in;
    if (t < valid) {
      const int64_t u = j.utt ? j.utt[i] : i;
      cast_row(j.src + (u * j.T_in + t) * j.ld_src, d, j.cols);
      if (j.ld_dst > j.cols) memset(d + j.cols, 0, (j.ld_dst - j.cols) * sizeof(uint16_t));
    } else {
      memset(d, 0, j.ld_dst * sizeof(uint16_t));
    }
  }
#ifdef HRIEMO_HAVE_AVX512
  if (g_avx512) _mm_sfence();   // non-temporal stores are globally visible before the worker reports the slab ready
#endif
}

}  // namespace

extern "C" int hriemo_host_pack_bf16(const float* src, int64_t ld_src, int64_t T_in, int64_t cols,
                                     const int32_t* utt, const int32_t* lens, void* dst_bf16, int64_t ld_dst,
                                     int64_t T_out, int64_t n, int64_t n_src, int32_t n_threads) {
  if (!src || !dst_bf16 || ld_src < cols || ld_dst < cols || T_in <= 0 || T_out <= 0 || cols <= 0 || n < 0)
    return hriemo::set_error(HRIEMO_ERR_INVALID, "host_pack_bf16: bad argument (n=%lld T_in=%lld T_out=%lld cols=%lld)",
                             (long long)n, (long long)T_in, (long long)T_out, (long long)cols);
  if (n == 0) return HRIEMO_OK;
  if (utt == nullptr && n > n_src)
    return hriemo::set_error(HRIEMO_ERR_INVALID, "host_pack_bf16: n=%lld exceeds the source batch (%lld)", (long long)n,
                             (long long)n_src);
  if (utt != nullptr)
    for (int64_t i = 0; i < n; ++i)
      if (utt[i] < 0 || utt[i] >= n_src)
        return hriemo::set_error(HRIEMO_ERR_INVALID, "host_pack_bf16: utt[%lld]=%d outside the source batch (%lld)",
                                 (long long)i, utt[i], (long long)n_src);
  PackJob j{src, ld_src, T_in, cols, utt, lens, static_cast<uint16_t*>(dst_bf16), ld_dst, T_out, n};
  const int64_t rows = n * T_out;
  int64_t nt = std::max<int64_t>(1, std::min<int64_t>(n_threads, rows / 256 + 1));
  if (nt == 1) {
    pack_rows(j, 0, rows);
    return HRIEMO_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  const int64_t per = (rows + nt - 1) / nt;
  for (int64_t k = 1; k < nt; ++k) {
    const int64_t r0 = k * per, r1 = std::min(rows, r0 + per);
    if (r0 < r1) pool.emplace_back([&j, r0, r1] { pack_rows(j, r0, r1); });
  }
  pack_rows(j, 0, std::min(rows, per));
  for (auto& th : pool) th.join();
  return HRIEMO_OK;
}
