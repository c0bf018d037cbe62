// HOST-side staging helper (no device work): gathers utterances out of the padded fp32 batch the
// reference's collate produced (scripts/fusion/train_fusion_seq_level_decoder.py:191-232) and writes
// them as bf16, trimmed to the slab's own length, into (pinned) staging memory -- half the bytes cross
// PCIe, and none of the padding beyond the slab maximum.  Plain C++ threads; round-to-nearest-even
// exactly like the GPU cast (cvt.rn.bf16x2.f32), so a host-packed slab is bit-identical to a
// device-cast one for every finite input.
#include <stdint.h>
#include <string.h>

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
void cast_row(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t u = s[i];
    const uint32_t r = (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
    dst[i] = ((u & 0x7fffffffu) > 0x7f800000u) ? static_cast<uint16_t>(0x7fffu) : static_cast<uint16_t>(r);
  }
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
}

}  // namespace

extern "C" int hriemo_host_pack_bf16(const float* src, int64_t ld_src, int64_t T_in, int64_t cols,
                                     const int32_t* utt, const int32_t* lens, void* dst_bf16, int64_t ld_dst,
                                     int64_t T_out, int64_t n, int32_t n_threads) {
  if (!src || !dst_bf16 || ld_src < cols || ld_dst < cols || T_in <= 0 || T_out <= 0 || cols <= 0 || n < 0)
    return hriemo::set_error(HRIEMO_ERR_INVALID, "host_pack_bf16: bad argument (n=%lld T_in=%lld T_out=%lld cols=%lld)",
                             (long long)n, (long long)T_in, (long long)T_out, (long long)cols);
  if (n == 0) return HRIEMO_OK;
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
