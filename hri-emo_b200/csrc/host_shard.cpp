// Packed feature shards: the on-disk side of the step before the path.
//
// The reference stores one torch pickle per utterance and modality,
//   {"hidden": [L, d] float32, "attention_mask": [L] long (1 = valid)}
// (scripts/iemocap_feature_extraction_seq_level/extract_audio_feats_wavlm_seq.py:118-135, loaded one by
// one at scripts/fusion/train_fusion_seq_level_decoder.py:139-156), which caps a data loader far below
// what one B200 consumes.  A shard holds many utterances in ONE file, laid out so that a slab of
// utterances can go from the page cache to pinned staging memory with plain row copies:
//
//   header (128 B)   magic "HRIEMOS1", version, dtype (1 = bf16, 2 = f32), n_utt, d_a, d_t, total rows,
//                    section offsets, max lengths
//   index            n_utt x {u64 row_a, u64 row_t, u32 len_a, u32 len_t, u64 reserved}
//   audio rows       [rows_a, d_a]  -- each utterance's rows up to its last valid position, back to back
//   text rows        [rows_t, d_t]
//   audio PAD bytes  [rows_a]       -- 1 = PAD (holes inside an utterance survive), like key_padding_mask
//   text PAD bytes   [rows_t]
//   meta             JSON (uids, labels, provenance), opaque to this reader
//
// Sections start on 4 KiB boundaries.  The writer is hri-emo_b200/hriemo/shards.py; this file is the
// reader: mmap + multi-threaded row copies into caller-provided (pinned) buffers, zero-padding each
// utterance to the slab's extents and producing the True = PAD masks of the reference's collate
// (scripts/fusion/train_fusion_seq_level_decoder.py:191-232).  HOST code only, no CUDA.
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/hriemo.h"

namespace hriemo {
int set_error(int code, const char* fmt, ...);   // host_common.cu
}

namespace {

struct ShardHeader {
  char magic[8];
  uint32_t version;
  uint32_t dtype;
  uint64_t n_utt;
  uint32_t d_a, d_t;
  uint64_t rows_a, rows_t;
  uint64_t off_index, off_audio, off_text, off_mask_a, off_mask_t, off_meta;
  uint64_t file_bytes;
  uint32_t max_len_a, max_len_t;
  uint64_t meta_bytes;
  uint8_t reserved[8];
};
static_assert(sizeof(ShardHeader) == 128, "shard header is 128 bytes");

struct ShardIndex {
  uint64_t row_a, row_t;
  uint32_t len_a, len_t;
  uint64_t reserved;
};
static_assert(sizeof(ShardIndex) == 32, "shard index entry is 32 bytes");

struct Shard {
  int fd;
  const uint8_t* base;
  size_t bytes;
  ShardHeader h;
  const ShardIndex* index;
  size_t elem;   // bytes per feature element
};

bool section_ok(const Shard& s, uint64_t off, uint64_t len) { return off <= s.bytes && len <= s.bytes - off; }

struct ReadJob {
  const Shard* s;
  const int64_t* utt;
  int64_t first, n;
  int64_t T_a, T_t;
  uint8_t* dst_a; uint8_t* dst_t; uint8_t* mask_a; uint8_t* mask_t;
};

// one modality of utterances [i0, i1) of the slab
void read_range(const ReadJob& j, bool audio, int64_t i0, int64_t i1) {
  const Shard& s = *j.s;
  const int64_t T = audio ? j.T_a : j.T_t;
  const size_t row_bytes = (audio ? s.h.d_a : s.h.d_t) * s.elem;
  const uint8_t* rows = s.base + (audio ? s.h.off_audio : s.h.off_text);
  const uint8_t* pads = s.base + (audio ? s.h.off_mask_a : s.h.off_mask_t);
  uint8_t* dst = audio ? j.dst_a : j.dst_t;
  uint8_t* msk = audio ? j.mask_a : j.mask_t;
  for (int64_t i = i0; i < i1; ++i) {
    const ShardIndex& e = s.index[j.utt ? j.utt[i] : j.first + i];
    const int64_t len = std::min<int64_t>(audio ? e.len_a : e.len_t, T);
    const uint64_t row0 = audio ? e.row_a : e.row_t;
    if (dst) {
      uint8_t* d = dst + static_cast<size_t>(i) * T * row_bytes;
      memcpy(d, rows + row0 * row_bytes, static_cast<size_t>(len) * row_bytes);
      memset(d + static_cast<size_t>(len) * row_bytes, 0, static_cast<size_t>(T - len) * row_bytes);
    }
    if (msk) {
      uint8_t* m = msk + static_cast<size_t>(i) * T;
      memcpy(m, pads + row0, static_cast<size_t>(len));
      memset(m + len, 1, static_cast<size_t>(T - len));
    }
  }
}

}  // namespace

extern "C" int hriemo_shard_open(const char* path, void** handle) {
  if (!path || !handle) return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_open: null argument");
  *handle = nullptr;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_open: cannot open %s: %s", path, strerror(errno));
  struct stat st;
  if (fstat(fd, &st) != 0 || static_cast<size_t>(st.st_size) < sizeof(ShardHeader)) {
    close(fd);
    return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_open: %s is too small to be a shard", path);
  }
  void* m = mmap(nullptr, st.st_size, PROT_READ, MAP_SHARED, fd, 0);
  if (m == MAP_FAILED) {
    close(fd);
    return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_open: mmap of %s failed: %s", path, strerror(errno));
  }
  Shard* s = new Shard();
  s->fd = fd;
  s->base = static_cast<const uint8_t*>(m);
  s->bytes = st.st_size;
  memcpy(&s->h, m, sizeof(ShardHeader));
  const ShardHeader& h = s->h;
  s->elem = h.dtype == 1 ? 2 : 4;
  // every size below is a product / sum of 64-bit fields a corrupt or crafted file controls: overflow-checked, so a
  // wrapped value can never pass the bounds test and send read_range outside the mapping
  uint64_t sz_index = 0, sz_audio = 0, sz_text = 0;
  bool ok = memcmp(h.magic, "HRIEMOS1", 8) == 0 && h.version == 1 && (h.dtype == 1 || h.dtype == 2) && h.d_a > 0 &&
            h.d_t > 0 && h.file_bytes == s->bytes &&
            !__builtin_mul_overflow(h.n_utt, static_cast<uint64_t>(sizeof(ShardIndex)), &sz_index) &&
            !__builtin_mul_overflow(h.rows_a, static_cast<uint64_t>(h.d_a), &sz_audio) &&
            !__builtin_mul_overflow(sz_audio, static_cast<uint64_t>(s->elem), &sz_audio) &&
            !__builtin_mul_overflow(h.rows_t, static_cast<uint64_t>(h.d_t), &sz_text) &&
            !__builtin_mul_overflow(sz_text, static_cast<uint64_t>(s->elem), &sz_text) &&
            section_ok(*s, h.off_index, sz_index) && section_ok(*s, h.off_audio, sz_audio) &&
            section_ok(*s, h.off_text, sz_text) && section_ok(*s, h.off_mask_a, h.rows_a) &&
            section_ok(*s, h.off_mask_t, h.rows_t) && section_ok(*s, h.off_meta, h.meta_bytes);
  if (ok) {
    s->index = reinterpret_cast<const ShardIndex*>(s->base + h.off_index);
    for (uint64_t i = 0; i < h.n_utt && ok; ++i) {
      const ShardIndex& e = s->index[i];
      // len <= rows first, then row <= rows - len: no sum that could wrap
      ok = e.len_a <= h.rows_a && e.row_a <= h.rows_a - e.len_a && e.len_t <= h.rows_t && e.row_t <= h.rows_t - e.len_t &&
           e.len_a <= h.max_len_a && e.len_t <= h.max_len_t;
    }
  }
  if (!ok) {
    munmap(m, st.st_size);
    close(fd);
    delete s;
    return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_open: %s is not a valid HRIEMOS1 shard (bad header, index or size)", path);
  }
  madvise(m, st.st_size, MADV_WILLNEED);
  *handle = s;
  return HRIEMO_OK;
}

extern "C" int hriemo_shard_info(void* handle, hriemo_shard_info_t* info) {
  if (!handle || !info) return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_info: null argument");
  const Shard* s = static_cast<const Shard*>(handle);
  info->n_utt = static_cast<int64_t>(s->h.n_utt);
  info->rows_a = static_cast<int64_t>(s->h.rows_a);
  info->rows_t = static_cast<int64_t>(s->h.rows_t);
  info->meta_bytes = static_cast<int64_t>(s->h.meta_bytes);
  info->d_a = static_cast<int32_t>(s->h.d_a);
  info->d_t = static_cast<int32_t>(s->h.d_t);
  info->dtype = static_cast<int32_t>(s->h.dtype);
  info->max_len_a = static_cast<int32_t>(s->h.max_len_a);
  info->max_len_t = static_cast<int32_t>(s->h.max_len_t);
  return HRIEMO_OK;
}

extern "C" int hriemo_shard_lengths(void* handle, int32_t* len_a, int32_t* len_t) {
  if (!handle || !len_a || !len_t) return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_lengths: null argument");
  const Shard* s = static_cast<const Shard*>(handle);
  for (uint64_t i = 0; i < s->h.n_utt; ++i) {
    len_a[i] = static_cast<int32_t>(s->index[i].len_a);
    len_t[i] = static_cast<int32_t>(s->index[i].len_t);
  }
  return HRIEMO_OK;
}

extern "C" int hriemo_shard_meta(void* handle, char* dst, int64_t dst_bytes) {
  if (!handle || !dst) return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_meta: null argument");
  const Shard* s = static_cast<const Shard*>(handle);
  if (dst_bytes < static_cast<int64_t>(s->h.meta_bytes))
    return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_meta: buffer of %lld bytes, need %llu", (long long)dst_bytes,
                             (unsigned long long)s->h.meta_bytes);
  memcpy(dst, s->base + s->h.off_meta, s->h.meta_bytes);
  return HRIEMO_OK;
}

extern "C" int hriemo_shard_read(void* handle, const int64_t* utt, int64_t first, int64_t n, int32_t T_a, int32_t T_t,
                                 void* dst_a, void* dst_t, uint8_t* mask_a, uint8_t* mask_t, int32_t n_threads) {
  if (!handle) return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_read: null handle");
  const Shard* s = static_cast<const Shard*>(handle);
  if (n < 0 || T_a <= 0 || T_t <= 0 || (!utt && (first < 0 || first + n > static_cast<int64_t>(s->h.n_utt))))
    return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_read: bad range first=%lld n=%lld of %llu (T_a=%d T_t=%d)",
                             (long long)first, (long long)n, (unsigned long long)s->h.n_utt, T_a, T_t);
  if (utt)
    for (int64_t i = 0; i < n; ++i)
      if (utt[i] < 0 || utt[i] >= static_cast<int64_t>(s->h.n_utt))
        return hriemo::set_error(HRIEMO_ERR_INVALID, "shard_read: utterance index %lld out of range", (long long)utt[i]);
  if (n == 0) return HRIEMO_OK;
  ReadJob j{s, utt, first, n, T_a, T_t, static_cast<uint8_t*>(dst_a), static_cast<uint8_t*>(dst_t), mask_a, mask_t};
  const int64_t nt = std::max<int64_t>(1, std::min<int64_t>(n_threads, n));
  auto work = [&j, n, nt](int64_t k) {
    const int64_t per = (n + nt - 1) / nt, i0 = k * per, i1 = std::min(n, i0 + per);
    if (i0 < i1) {
      read_range(j, true, i0, i1);
      read_range(j, false, i0, i1);
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  for (int64_t k = 1; k < nt; ++k) pool.emplace_back(work, k);
  work(0);
  for (auto& th : pool) th.join();
  return HRIEMO_OK;
}

extern "C" int hriemo_shard_close(void* handle) {
  if (!handle) return HRIEMO_OK;
  Shard* s = static_cast<Shard*>(handle);
  munmap(const_cast<uint8_t*>(s->base), s->bytes);
  close(s->fd);
  delete s;
  return HRIEMO_OK;
}
