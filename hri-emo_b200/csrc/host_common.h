// Host-side helpers shared by the C-ABI entry points: error text, launch
// counter, TMA tensor-map encoding through the driver entry point (no -lcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hriemo.h"

namespace hriemo {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

// 2-D bf16 tensor map: dim0 (contiguous) extent `inner`, dim1 extent `outer`,
// row pitch `pitch_elems`; box = box_inner x box_outer, 128-byte swizzle (or 64-byte with
// swizzle_bytes = 64), out-of-bounds elements read as zero.  Returns 0 or a negative error code.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes = 128);

// 3-D bf16 tensor map: dims {d0 (contiguous), d1, d2}, pitches in elements for d1, d2; no swizzle
// unless swizzle_bytes is 64 or 128.  Out-of-bounds elements read as zero / are not written.
int make_tmap_bf16_3d_plain(CUtensorMap* map, const void* base, const uint64_t dims[3],
                            const uint64_t pitch_elems[2], const uint32_t box[3], int swizzle_bytes = 0);

int device_sm_count();

// Function attributes (opt-in shared memory) are per device: `done` is a per-kernel bitmask of the
// devices the attribute has been set on; returns true when the current device still needs it.
bool device_needs_attr(uint64_t* done);

#define HRIEMO_REQUIRE(cond, ...)                                      \
  do {                                                                 \
    if (!(cond)) return ::hriemo::set_error(HRIEMO_ERR_INVALID, __VA_ARGS__); \
  } while (0)

}  // namespace hriemo
