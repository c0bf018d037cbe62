// Weight / bias gradient of y = x W^T + b on sm_100a tensor cores (the bulk of the
// backward FLOPs):   dW[N,K] = dY[M,N]^T . X[M,K]     db[N] = sum_m dY[m, n]
// (what loss.backward() accumulates into nn.Linear.weight.grad / .bias.grad at
// scripts/fusion/train_fusion_seq_level_decoder.py:332 for every projection and FFN layer of models/*.py).
//
// The contraction runs over the ROWS of both operands (M = B*T, about a million), so neither operand is
// K-major: both are consumed exactly as they lie in HBM, as MN-major UMMA operands -- the form the attention
// kernel uses for V -- instead of being transposed first.  One CTA owns a 128 (n) x BNK (k) tile of dW and one
// of `splits` ranges of m:
//   warp 0      TMA producer: per 64-row block of m, 2 boxes of dY [64 m][64 n] and BNK/64 boxes of X
//               [64 m][64 k], 128-byte-swizzled, into a 4-stage ring;
//   warp 1      tcgen05.mma issuer (one thread): D[128 x BNK] += A[128 x 16] B[16 x BNK], four per block, fp32
//               accumulation in tensor memory for the whole m range;
//   warps 2..5  epilogue: tcgen05.ld, fp32 partial tile to workspace[split][n][k].
// A second kernel adds the `splits` partial tiles in a fixed order (deterministic, unlike atomics) into dW,
// optionally on top of what is there (gradient accumulation).  db rides along: the epilogue warps of the first k tile's
// CTAs sum the columns of the dY stages while the MMAs run (see the kernel), and the same second kernel adds the partials.
#include <stdlib.h>

#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

constexpr int WG_BN = 128;      // n rows of dW per tile (UMMA M)
constexpr int WG_MB = 64;       // rows of m per pipeline stage (four K=16 MMA steps)
constexpr int WG_STAGES = 4;
constexpr int WG_GROUP = WG_MB * 128;  // one [64 m][64 columns] box: 128 B per row, 128B swizzle (with 64-byte rows the
                                       // operand reads ran at half rate: ~190 cycles per 128-cycle MMA)
constexpr int WG_THREADS = 192;
constexpr uint32_t kUmmaAMajorMN = 1u << 15;   // instruction-descriptor flag: A operand is MN-major

// db[n] = sum_m dY[m, n] rides along with dW: the four epilogue warps have nothing to do until the accumulator is
// complete, and the dY stage [64 m][128 n] is in shared memory anyway.  The CTAs (or pairs) of the K / 256 k tiles of one
// (n tile, row range) stream the SAME dY blocks, so they deal them out round-robin: the CTA of k tile kt sums blocks
// i = kt, kt + k_tiles, ...  -- one k tile doing all of it made its CTAs the stragglers of the launch once the main loop
// ran at full rate.  Thread (half, pair) adds rows [32 half, 32 half + 32) of columns 2 pair, 2 pair + 1 of a stage: one
// 4-byte load per row, a warp reads one 128-byte row of a column group (32 distinct banks), four independent sums.
// The warps follow EVERY stage's full barrier in order (an mbarrier wait only sees a parity: skipping uses would alias),
// read the stages that are theirs and tell the producer through read_done, which it awaits before refilling such a stage.
// One fp32 partial per (split, k tile, half) and column, added in a fixed order by sum_partials_kernel like the dW tiles.
template <int STAGES>
__device__ __forceinline__ void wgrad_colsum_rows(float* __restrict__ colsum, uint32_t sA, int a_stage_bytes, uint32_t bar_wait0,
                                                  uint32_t bar_read_done, int nblk, int kt, int k_tiles, int split, int N, int n0,
                                                  int warp, int lane) {
  const int et = (warp - 2) * 32 + lane;
  const int half = et >> 6, n_in = (et & 63) * 2;
  const uint32_t grp = static_cast<uint32_t>(n_in >> 6), chunk = static_cast<uint32_t>((n_in & 63) >> 3);
  const uint32_t within = static_cast<uint32_t>(n_in & 7) * 2u;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  for (int i = 0; i < nblk; ++i) {
    const uint32_t s = i % STAGES, par = (i / STAGES) & 1u;
    mbar_wait(bar_wait0 + s * 8, par);
    if (i % k_tiles != kt) continue;
    const uint32_t a = sA + s * a_stage_bytes + grp * WG_GROUP;
#pragma unroll 8
    for (int r = 0; r < 32; r += 2) {
      const uint32_t m = static_cast<uint32_t>(half * 32 + r);
      uint32_t w0, w1;   // 128-byte rows, SWIZZLE_128B: the 16-byte chunk index is XORed with the row index mod 8
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(a + m * 128u + ((chunk ^ (m & 7u)) << 4) + within));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(a + (m + 1u) * 128u + ((chunk ^ ((m + 1u) & 7u)) << 4) + within));
      s0 += bf16_lo(w0);
      s1 += bf16_hi(w0);
      s2 += bf16_lo(w1);
      s3 += bf16_hi(w1);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_read_done + s * 8);
  }
  float* dst = colsum + static_cast<int64_t>((split * k_tiles + kt) * 2 + half) * N + n0 + n_in;
  dst[0] = s0 + s2;
  dst[1] = s1 + s3;
}
// the producer's side of read_done: before stage s is refilled for block i, the column-sum warps must have read block
// i - STAGES if it was one of theirs; `parity_bits` holds, per stage, the parity of the next read_done phase
template <int STAGES>
__device__ __forceinline__ void wgrad_await_readers(uint32_t bar_read_done, uint32_t& parity_bits, int i, int kt, int k_tiles) {
  if (i >= STAGES && (i - STAGES) % k_tiles == kt) {
    const uint32_t s = i % STAGES;
    mbar_wait(bar_read_done + s * 8, (parity_bits >> s) & 1u);
    parity_bits ^= 1u << s;
  }
}

template <int BNK>
struct WgradSmem {
  static constexpr int A_STAGE = (WG_BN / 64) * WG_GROUP;   // 16 KB
  static constexpr int B_STAGE = (BNK / 64) * WG_GROUP;     // 32 KB (BNK = 256) / 16 KB (128)
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = WG_STAGES * A_STAGE;
  static constexpr int BAR_OFF = B_OFF + WG_STAGES * B_STAGE;   // full[4] empty[4] read_done[4] tfull
  static constexpr int TMEM_SLOT_OFF = BAR_OFF + (3 * WG_STAGES + 1) * 8;
  static constexpr int DYN_BYTES = TMEM_SLOT_OFF + 16 + 1024;
};

template <int BNK>
__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                  float* __restrict__ partials, float* __restrict__ colsum, int64_t M, int N, int K, int blocks_per_split) {
  using L = WgradSmem<BNK>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base + L::A_OFF, sB = base + L::B_OFF;
  const uint32_t bar_full = base + L::BAR_OFF, bar_empty = bar_full + WG_STAGES * 8, bar_rdone = bar_empty + WG_STAGES * 8;
  const uint32_t bar_tfull = bar_rdone + WG_STAGES * 8;
  const uint32_t tmem_slot = base + L::TMEM_SLOT_OFF;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * BNK, n0 = blockIdx.y * WG_BN, split = blockIdx.z;
  const int total_blocks = static_cast<int>((M + WG_MB - 1) / WG_MB);
  const int blk0 = split * blocks_per_split;
  const int blk1 = min(total_blocks, blk0 + blocks_per_split);
  const int nblk = max(0, blk1 - blk0);
  const bool with_colsum = colsum != nullptr;   // db: see wgrad_colsum_rows
  const int kt = blockIdx.x, k_tiles = gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
      mbar_init(bar_rdone + s * 8, 4);   // the four column-sum warps
    }
    mbar_init(bar_tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BNK>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // The single TMA / MMA thread is chosen with elect.sync, not `lane == 0`: cp.async.bulk.tensor and tcgen05.mma take
  // uniform-register operands, and only behind elect.sync does the compiler know one lane is active -- behind
  // `if (lane == 0)` every one of the 12 loads and 4 MMAs of a stage was wrapped in an ELECT / R2UR.BROADCAST / branch-back
  // loop (the stall samples of profiles/r01_ncu_wgrad_v13_summary.txt), as in the forward kernels before round 2.
  if (warp == 0) {
    if (elect_one()) {
      uint32_t rd_par = 0;
      for (int i = 0; i < nblk; ++i) {
        const uint32_t s = i % WG_STAGES, par = (i / WG_STAGES) & 1u;
        mbar_wait(bar_empty + s * 8, par ^ 1);
        if (with_colsum) wgrad_await_readers<WG_STAGES>(bar_rdone, rd_par, i, kt, k_tiles);
        mbar_arrive_expect_tx(bar_full + s * 8, L::A_STAGE + L::B_STAGE);
        const int m = (blk0 + i) * WG_MB;   // rows past M read as zero (no contribution)
        for (int c = 0; c < WG_BN / 64; ++c)
          tma_load_2d(&tm_dy, bar_full + s * 8, sA + s * L::A_STAGE + c * WG_GROUP, n0 + c * 64, m);
        for (int c = 0; c < BNK / 64; ++c)
          tma_load_2d(&tm_x, bar_full + s * 8, sB + s * L::B_STAGE + c * WG_GROUP, k0 + c * 64, m);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(WG_BN, BNK) | kUmmaAMajorMN | kUmmaBMajorMN;
      for (int i = 0; i < nblk; ++i) {
        const uint32_t s = i % WG_STAGES, par = (i / WG_STAGES) & 1u;
        mbar_wait(bar_full + s * 8, par);
        tc_fence_after_sync();
        const uint64_t a_desc = umma_desc_mn_sw128(sA + s * L::A_STAGE, WG_GROUP);
        const uint64_t b_desc = umma_desc_mn_sw128(sB + s * L::B_STAGE, WG_GROUP);
#pragma unroll
        for (int st = 0; st < WG_MB / 16; ++st)   // 16 rows of m = 16 x 128 B inside every column group
          umma_bf16(tmem_base, a_desc + ((st * 16 * 128) >> 4), b_desc + ((st * 16 * 128) >> 4), idesc, (i | st) != 0);
        umma_commit(bar_empty + s * 8);
      }
      if (nblk > 0) umma_commit(bar_tfull);
      else mbar_arrive(bar_tfull);   // an empty split still hands a (zero) tile to the epilogue
    }
  } else {
    if (with_colsum)
      wgrad_colsum_rows<WG_STAGES>(colsum, sA, L::A_STAGE, bar_full, bar_rdone, nblk, kt, k_tiles, split, N, n0, warp, lane);
    // epilogue: this thread's row of the tile is its TMEM lane
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    mbar_wait(bar_tfull, 0);
    tc_fence_after_sync();
    float* dst = partials + (static_cast<int64_t>(split) * N + n0 + row) * K + k0;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BNK / 32; ++c) {
      uint32_t v[32];
      if (nblk > 0) {
        tmem_ld32(t_row + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (n0 + row < N) {
        float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          d4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                              __uint_as_float(v[4 * i + 3]));
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<BNK>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair form (cluster of two SMs, tcgen05 cta_group::2): the pair owns a 256 (n) x 256 (k) tile of dW.  Each CTA loads
// ITS 128 columns of dY and ITS 128 columns of X per 64-row block (32 KB per stage instead of the 48 KB of the single-CTA
// 128 x 256 tile: 131 instead of 87 FLOP per byte brought into shared memory -- the single-CTA form runs ~1 000 cycles
// per stage for 512 cycles of tensor work, bound by what one SM can pull from L2), the leader issues M = 256 MMAs that
// read both CTAs' shared memory, and each CTA's tensor memory receives its 128 rows x 256 columns of the tile.
//   both CTAs  warp 0: TMA producer of the CTA's own halves, completing on the LEADER's full barrier
//   leader     warp 1: tcgen05.mma.cta_group::2 issuer; its commits free the stage in both CTAs (multicast)
//   both CTAs  warps 2..5: column sums of the CTA's dY half (its share of the blocks), then the epilogue of the CTA's rows.
// The column-sum warps of the peer cannot wait on the leader's full barrier, so the issuer relays each "stage landed"
// to a second barrier (full2) in both CTAs (release / acquire at cluster scope).
constexpr int WGP_STAGES = 6;

struct WgradPairSmem {
  static constexpr int A_STAGE = 2 * WG_GROUP;   // [64 m][128 n], 16 KB
  static constexpr int B_STAGE = 2 * WG_GROUP;   // [64 m][128 k], 16 KB
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = WGP_STAGES * A_STAGE;
  static constexpr int BAR_OFF = B_OFF + WGP_STAGES * B_STAGE;   // full[S] empty[S] full2[S] read_done[S] tfull
  static constexpr int TMEM_SLOT_OFF = BAR_OFF + (4 * WGP_STAGES + 1) * 8;
  static constexpr int DYN_BYTES = TMEM_SLOT_OFF + 16 + 1024;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_wgrad_pair_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                       float* __restrict__ partials, float* __restrict__ colsum, int64_t M, int N, int K,
                       int blocks_per_split, int k_tiles) {
  using L = WgradPairSmem;
  constexpr int S = WGP_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base + L::A_OFF, sB = base + L::B_OFF;
  const uint32_t bar_full = base + L::BAR_OFF, bar_empty = bar_full + S * 8, bar_full2 = bar_empty + S * 8;
  const uint32_t bar_rdone = bar_full2 + S * 8, bar_tfull = bar_rdone + S * 8;
  const uint32_t tmem_slot = base + L::TMEM_SLOT_OFF;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int kt = pair % k_tiles, nt = pair / k_tiles;
  const int k_tile0 = kt * 256;                               // the pair's 256 columns of dW
  const int n0 = nt * 256 + static_cast<int>(rank) * 128;     // this CTA's 128 rows of dW
  const int split = blockIdx.z;
  const int total_blocks = static_cast<int>((M + WG_MB - 1) / WG_MB);
  const int blk0 = split * blocks_per_split;
  const int blk1 = min(total_blocks, blk0 + blocks_per_split);
  const int nblk = max(0, blk1 - blk0);
  const bool with_colsum = colsum != nullptr;   // db: see wgrad_colsum_rows

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
      mbar_init(bar_full2 + s * 8, 1);
      mbar_init(bar_rdone + s * 8, 4);   // the four column-sum warps
    }
    mbar_init(bar_tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<256>(tmem_slot);
  tc_fence_before_sync();
  cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / TMA completion
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t rd_par = 0;
      for (int i = 0; i < nblk; ++i) {
        const uint32_t s = i % S, par = (i / S) & 1u;
        mbar_wait(bar_empty + s * 8, par ^ 1);   // this CTA's slot is free
        if (with_colsum) wgrad_await_readers<S>(bar_rdone, rd_par, i, kt, k_tiles);
        if (leader) mbar_arrive_expect_tx(bar_full + s * 8, 2 * (L::A_STAGE + L::B_STAGE));   // both CTAs' loads
        const int m = (blk0 + i) * WG_MB;        // rows past M read as zero (no contribution)
        for (int c = 0; c < 2; ++c)
          tma_load_2d_pair(&tm_dy, bar_full + s * 8, sA + s * L::A_STAGE + c * WG_GROUP, n0 + c * 64, m);
        for (int c = 0; c < 2; ++c)
          tma_load_2d_pair(&tm_x, bar_full + s * 8, sB + s * L::B_STAGE + c * WG_GROUP,
                           k_tile0 + static_cast<int>(rank) * 128 + c * 64, m);
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, 256) | kUmmaAMajorMN | kUmmaBMajorMN;
      const uint32_t peer_full2 = mapa_shared(bar_full2, 1u), peer_tfull = mapa_shared(bar_tfull, 1u);
      for (int i = 0; i < nblk; ++i) {
        const uint32_t s = i % S, par = (i / S) & 1u;
        mbar_wait(bar_full + s * 8, par);
        if (with_colsum) {
          // relaxed: the stage's bytes were in each CTA's shared memory before their completion reached the full barrier
          // this thread has just seen; a release at cluster scope here is a full fence per stage in the issuer (2 x slower)
          mbar_arrive(bar_full2 + s * 8);
          mbar_arrive_cluster(peer_full2 + s * 8);
        }
        tc_fence_after_sync();
        const uint64_t a_desc = umma_desc_mn_sw128(sA + s * L::A_STAGE, WG_GROUP);
        const uint64_t b_desc = umma_desc_mn_sw128(sB + s * L::B_STAGE, WG_GROUP);
#pragma unroll
        for (int st = 0; st < WG_MB / 16; ++st)
          umma_bf16_pair(tmem_base, a_desc + ((st * 16 * 128) >> 4), b_desc + ((st * 16 * 128) >> 4), idesc, (i | st) != 0);
        umma_commit_pair(bar_empty + s * 8, 3);
      }
      if (nblk > 0) {
        umma_commit_pair(bar_tfull, 3);
      } else {   // an empty split still hands a (zero) tile to both epilogues
        mbar_arrive(bar_tfull);
        mbar_arrive_cluster(peer_tfull);
      }
    }
  } else {
    if (with_colsum)
      wgrad_colsum_rows<S>(colsum, sA, L::A_STAGE, bar_full2, bar_rdone, nblk, kt, k_tiles, split, N, n0, warp, lane);
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    mbar_wait(bar_tfull, 0);
    tc_fence_after_sync();
    float* dst = partials + (static_cast<int64_t>(split) * N + n0 + row) * K + k_tile0;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < 256 / 32; ++c) {
      uint32_t v[32];
      if (nblk > 0) {
        tmem_ld32(t_row + c * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        d4[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                            __uint_as_float(v[4 * i + 3]));
    }
    tc_fence_before_sync();
  }
  tc_fence_before_sync();
  cluster_sync_all();   // neither CTA leaves (or frees tensor memory) while the other still uses the pair's resources
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_pair<256>(tmem_base);
  }
}

// out[i] = (accumulate ? out[i] : 0) + sum_s partials[s][i], s in a fixed order
__global__ void sum_partials_kernel(const float* __restrict__ partials, int splits, int64_t n, float* __restrict__ out,
                                    int accumulate) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc = accumulate ? out[i] : 0.0f;
    for (int s = 0; s < splits; ++s) acc += partials[static_cast<int64_t>(s) * n + i];
    out[i] = acc;
  }
}

// The dW and db sums of one wgrad call in ONE launch: the last ceil(n1 / 256) CTAs take the bias partials (one column per
// thread), the others walk the dW partials grid-stride.  (Two launches per call were ~50 launches of 7 - 14 us per step.)
__global__ void sum_partials2_kernel(const float* __restrict__ p0, int splits0, int64_t n0, float* __restrict__ out0,
                                     const float* __restrict__ p1, int splits1, int n1, float* __restrict__ out1,
                                     int accumulate) {
  const int ctas1 = (n1 + static_cast<int>(blockDim.x) - 1) / static_cast<int>(blockDim.x);
  const int ctas0 = static_cast<int>(gridDim.x) - ctas1;
  if (static_cast<int>(blockIdx.x) >= ctas0) {
    const int i = (static_cast<int>(blockIdx.x) - ctas0) * static_cast<int>(blockDim.x) + static_cast<int>(threadIdx.x);
    if (i >= n1) return;
    float acc = accumulate ? out1[i] : 0.0f;
    for (int s = 0; s < splits1; ++s) acc += p1[static_cast<int64_t>(s) * n1 + i];
    out1[i] = acc;
    return;
  }
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n0;
       i += static_cast<int64_t>(ctas0) * blockDim.x) {
    float acc = accumulate ? out0[i] : 0.0f;
    for (int s = 0; s < splits0; ++s) acc += p0[static_cast<int64_t>(s) * n0 + i];
    out0[i] = acc;
  }
}

// out[c][r] = in[r][c] for a bf16 matrix (weights: [N,K] -> [K,N], the operand of dX = dY . W)
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld_in, __nv_bfloat16* __restrict__ out,
                      int64_t ld_out, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (r0 + r < rows && c0 + tx < cols) tile[r][tx] = in[static_cast<int64_t>(r0 + r) * ld_in + c0 + tx];
  __syncthreads();
  for (int c = ty; c < 32; c += 8)
    if (c0 + c < cols && r0 + tx < rows) out[static_cast<int64_t>(c0 + c) * ld_out + r0 + tx] = tile[tx][c];
}

// Number of row ranges: the CTAs (tiles x splits) should fill whole waves of the SMs -- 360 CTAs on 148 SMs run three
// rounds for 2.4 rounds of work -- with at least 8 blocks (512 rows) per split and as few splits as that allows
// (every split costs a partial tile in the workspace and in the final sum).
static bool wgrad_use_pairs(int64_t M, int N, int K) {
  static const int env = [] {
    const char* e = getenv("HRIEMO_WGRAD_PAIR");   // 0 / 1: A/B measurements in one build
    return e == nullptr ? -1 : atoi(e);
  }();
  if (N % 256 != 0 || K % 256 != 0 || env == 0) return false;
  (void)M;
  return true;
}
static int wgrad_splits(int64_t M, int N, int K, int bnk) {
  const bool pairs = bnk == 256 && wgrad_use_pairs(M, N, K);
  // units that own a tile: CTAs with a 128 x bnk tile, or CTA pairs with a 256 x 256 one
  const int64_t tiles = pairs ? static_cast<int64_t>(N / 256) * (K / 256) : static_cast<int64_t>(N / WG_BN) * (K / bnk);
  const int64_t blocks = (M + WG_MB - 1) / WG_MB;
  const int64_t sms = pairs ? device_sm_count() / 2 : device_sm_count();
  int64_t s_max = blocks / 8;
  if (s_max > 64) s_max = 64;
  if (s_max < 1) s_max = 1;
  int best = 1;
  double best_score = -1.0;
  for (int64_t sp = 1; sp <= s_max; ++sp) {
    const int64_t ctas = tiles * sp;
    const int64_t rounds = (ctas + sms - 1) / sms;
    double eff = static_cast<double>(ctas) / static_cast<double>(rounds * sms);   // filled fraction of the rounds
    if (rounds > 4) break;                                                         // enough parallelism: stop growing
    const double score = eff - 0.004 * static_cast<double>(sp);
    if (score > best_score) { best_score = score; best = static_cast<int>(sp); }
  }
  return best;
}
// workspace rows of the bias partials: one per (split, k tile, half)
static int64_t wgrad_colsum_rows_count(int splits, int K, int bnk) { return 2ll * splits * (K / bnk); }

template <int BNK>
static int launch_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, int64_t M, int N, int K,
                        float* partials, float* colsum, int splits, cudaStream_t stream) {
  using L = WgradSmem<BNK>;
  CUtensorMap tm_dy, tm_x;
  int rc = make_tmap_bf16_2d(&tm_dy, dY, (uint64_t)N, (uint64_t)M, (uint64_t)lddy, 64, WG_MB, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_x, X, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, 64, WG_MB, 128);
  if (rc) return rc;
  static uint64_t attr_done = 0;
  if (device_needs_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel<BNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "linear_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const int64_t blocks = (M + WG_MB - 1) / WG_MB;
  const int bps = static_cast<int>((blocks + splits - 1) / splits);
  dim3 grid(K / BNK, N / WG_BN, splits);
  gemm_wgrad_kernel<BNK><<<grid, WG_THREADS, L::DYN_BYTES, stream>>>(tm_dy, tm_x, partials, colsum, M, N, K, bps);
  return check_launch("linear_wgrad");
}

static int launch_wgrad_pair(const void* dY, int64_t lddy, const void* X, int64_t ldx, int64_t M, int N, int K,
                             float* partials, float* colsum, int splits, cudaStream_t stream) {
  using L = WgradPairSmem;
  CUtensorMap tm_dy, tm_x;
  int rc = make_tmap_bf16_2d(&tm_dy, dY, (uint64_t)N, (uint64_t)M, (uint64_t)lddy, 64, WG_MB, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_x, X, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, 64, WG_MB, 128);
  if (rc) return rc;
  static uint64_t attr_done = 0;
  if (device_needs_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "linear_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const int64_t blocks = (M + WG_MB - 1) / WG_MB;
  const int bps = static_cast<int>((blocks + splits - 1) / splits);
  const int k_tiles = K / 256;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * k_tiles * (N / 256)), 1, static_cast<unsigned>(splits));
  cfg.blockDim = dim3(WG_THREADS);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_wgrad_pair_kernel, tm_dy, tm_x, partials, colsum, M, N, K, bps, k_tiles);
  if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "linear_wgrad: launch failed: %s", cudaGetErrorString(e));
  return check_launch("linear_wgrad");
}

}  // namespace hriemo

using namespace hriemo;

extern "C" int64_t hriemo_linear_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K) {
  if (M <= 0 || N <= 0 || K <= 0 || N % WG_BN != 0 || K % 128 != 0) return 0;
  const int splits = wgrad_splits(M, N, K, K % 256 == 0 ? 256 : 128);
  return (static_cast<int64_t>(splits) * N * K + wgrad_colsum_rows_count(splits, K, K % 256 == 0 ? 256 : 128) * N) * static_cast<int64_t>(sizeof(float));
}

extern "C" int hriemo_linear_wgrad_bf16(const void* dY, int64_t lddy, const void* X, int64_t ldx, int64_t M, int32_t N,
                                        int32_t K, float* dW, float* db, int32_t accumulate, void* workspace,
                                        void* stream) {
  HRIEMO_REQUIRE(dY && X && dW && workspace && M > 0, "linear_wgrad: bad argument");
  HRIEMO_REQUIRE(N > 0 && K > 0 && N % WG_BN == 0 && K % 128 == 0, "linear_wgrad: N=%d and K=%d must be multiples of 128", N, K);
  HRIEMO_REQUIRE(lddy % 8 == 0 && ldx % 8 == 0 && lddy >= N && ldx >= K &&
                     (reinterpret_cast<uintptr_t>(dY) & 15u) == 0 && (reinterpret_cast<uintptr_t>(X) & 15u) == 0 &&
                     (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0,
                 "linear_wgrad: operands must be 16-byte aligned with leading dimensions that are multiples of 8");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* partials = static_cast<float*>(workspace);
  const int bnk = K % 256 == 0 ? 256 : 128;
  const int splits = wgrad_splits(M, N, K, bnk);
  const int64_t nk = static_cast<int64_t>(N) * K;
  float* bpart = db != nullptr ? partials + static_cast<int64_t>(splits) * nk : nullptr;
  int rc = bnk == 256 && wgrad_use_pairs(M, N, K) ? launch_wgrad_pair(dY, lddy, X, ldx, M, N, K, partials, bpart, splits, s)
           : bnk == 256                         ? launch_wgrad<256>(dY, lddy, X, ldx, M, N, K, partials, bpart, splits, s)
                                                : launch_wgrad<128>(dY, lddy, X, ldx, M, N, K, partials, bpart, splits, s);
  if (rc) return rc;
  int64_t gb = (nk + 255) / 256;
  if (gb > static_cast<int64_t>(device_sm_count()) * 8) gb = static_cast<int64_t>(device_sm_count()) * 8;
  const unsigned g = static_cast<unsigned>(gb);
  if (db == nullptr) {
    sum_partials_kernel<<<g, 256, 0, s>>>(partials, splits, nk, dW, accumulate);
    return check_launch("linear_wgrad (reduce)");
  }
  sum_partials2_kernel<<<g + static_cast<unsigned>((N + 255) / 256), 256, 0, s>>>(
      partials, splits, nk, dW, bpart, static_cast<int>(wgrad_colsum_rows_count(splits, K, bnk)), N, db, accumulate);
  return check_launch("linear_wgrad (reduce, with bias)");
}

extern "C" int hriemo_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int32_t rows, int32_t cols,
                                     void* stream) {
  HRIEMO_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "transpose: bad argument");
  transpose_bf16_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), ld_in, static_cast<__nv_bfloat16*>(out), ld_out, rows, cols);
  return check_launch("transpose_bf16");
}
