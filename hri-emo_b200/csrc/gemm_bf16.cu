// Persistent warp-specialised bf16 GEMM for sm_100a:
//   out = epilogue(A[M,K] . W[N,K]^T + bias)
// TMA (128B-swizzled tiles) -> shared-memory ring -> tcgen05.mma (fp32 accumulators
// in TMEM, double-buffered) -> epilogue warps (tcgen05.ld, bias/ReLU/residual,
// bf16 tiles staged in swizzled shared memory + TMA store, fp32 direct stores).
//
// Replaces the nn.Linear / MHA projection calls of the reference forward
// (models/cross_modal_block_tacfn.py:24-52,74-119; models/emotion_decoder.py:14-27;
//  models/mosei_fusion_with_emotion_decoder.py:41-42).
//
// Roles (128 + 32*EW threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4..4+EW-1 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31, one accumulator
// row per thread; with EW = 8 two warps share a lane quadrant and split the tile's columns).
//
// CTAS = 2 is the CTA-pair form (cluster of two SMs, tcgen05 cta_group::2): the pair owns a
// 256 x BN output tile; each CTA loads its 128 rows of A and its BN/2 rows of W, the leader's
// single thread issues M=256 MMAs that read both CTAs' shared memory, and each CTA's tensor
// memory receives (and its epilogue warps drain) its own 128 rows.  Per SM this halves the
// W bytes that TMA writes into and the tensor core reads out of shared memory per FLOP, which is
// what bounds the one-CTA form (24 KB of smem traffic per 128-cycle K16 step).
#include "host_common.h"
#include "sm100_ptx.cuh"

#include <cstdlib>

namespace hriemo {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int gemm_threads(int epi_warps) { return 128 + 32 * epi_warps; }
constexpr int OUT_BOX_BYTES = 32 * 128;  // one epilogue warp's staging box: 32 rows x 64 bf16

struct GemmKernelParams {
  int64_t M;
  int N, K;
  int epilogue;
  const float* bias;
  void* out;
  int64_t ldo;
  const void* resid;
  int64_t ldr;
  // fused LayerNorm (see hriemo.h): A rows / residual rows are pre-LayerNorm tensors with known
  // per-row (mean, rstd); per-slab (sum, sum of squares) of the output rows are written out
  const float2* a_stats;
  const float* a_colsum;
  const float2* resid_stats;
  const float* resid_gamma;
  const float* resid_beta;
  float2* stats_out;
  int num_n_blocks, num_k_blocks;
  int64_t num_tiles;
  int contiguous_walk;
};

template <int BN, int STAGES, int CTAS, int EW>
struct GemmSmem {
  static constexpr int B_STAGE_BYTES = (BN / CTAS) * BK * 2;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * A_STAGE_BYTES;
  static constexpr int OUT_OFF = B_OFF + STAGES * B_STAGE_BYTES;   // EW warps x 2 boxes of [32][64] bf16
  // per-column epilogue vectors (bias | colsum or gamma | beta), copied once per CTA when they fit
  static constexpr int VEC_OFF = OUT_OFF + 2 * EW * OUT_BOX_BYTES;
  static constexpr int VEC_BYTES = 0;  // (a 32 KB region cost a pipeline stage: slower overall)
  static constexpr int BAR_OFF = VEC_OFF + VEC_BYTES;
  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], resid_full[EW warps][2], tmem_ptr
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4 + 2 * EW) * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
};

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&f)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 v;
    v.x = pack_bf16(f[i * 8 + 0], f[i * 8 + 1]);
    v.y = pack_bf16(f[i * 8 + 2], f[i * 8 + 3]);
    v.z = pack_bf16(f[i * 8 + 4], f[i * 8 + 5]);
    v.w = pack_bf16(f[i * 8 + 6], f[i * 8 + 7]);
    d4[i] = v;
  }
}

template <int BN, int STAGES, int CTAS, int EW>
__global__ void __launch_bounds__(gemm_threads(EW), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_r,
                 const GemmKernelParams p) {
  using L = GemmSmem<BN, STAGES, CTAS, EW>;
  static_assert(EW == 4 || EW == 8, "one or two epilogue warps per TMEM lane quadrant");
  constexpr int SLABS = (BN / 64) / (EW / 4);  // 64-column slabs each epilogue warp drains per tile
  constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulator stages (power of two: 256 or 512)
  constexpr uint32_t STAGE_TX = (A_STAGE_BYTES + L::B_STAGE_BYTES) * CTAS;  // both CTAs' loads land on the leader's barrier
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // tiles are owned by a CTA (CTAS = 1) or a CTA pair (CTAS = 2) and are BM*CTAS rows tall
  // tile walk of an owner (CTA or CTA pair): strided (owner, owner + owners, ...: the N tiles of an M block run on
  // neighbouring owners at about the same time and share its A rows through L2 -- when they stay in step) or, with
  // p.contiguous_walk, a contiguous range of the M-major tile order (an owner visits the N tiles of an M block one after the
  // other and re-reads its A rows from L2 by itself).
  const int64_t owners = gridDim.x / CTAS, owner = blockIdx.x / CTAS;
  const int64_t tile_first = p.contiguous_walk ? p.num_tiles * owner / owners : owner;
  const int64_t tile_end = p.contiguous_walk ? p.num_tiles * (owner + 1) / owners : p.num_tiles;
  const int64_t tile_step = p.contiguous_walk ? 1 : owners;
  constexpr int TILE_M = BM * CTAS;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base + L::A_OFF;
  const uint32_t sB = smem_base + L::B_OFF;
  const uint32_t sOut = smem_base + L::OUT_OFF;
  // Per-column vectors of the epilogue live in shared memory for the kernel's lifetime when
  // 3 x N floats fit (an L1 hit costs a long-scoreboard round trip per use; ld.shared does not).
  const bool vec_smem = L::VEC_BYTES > 0 && 3 * p.N * 4 <= L::VEC_BYTES;
  const uint32_t sVec = smem_base + L::VEC_OFF;
  if (vec_smem) {
    float* vs = reinterpret_cast<float*>(smem_raw + (sVec - smem_u32(smem_raw)));
    const float* v1 = p.a_colsum != nullptr ? p.a_colsum : p.resid_gamma;
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
      vs[i] = p.bias != nullptr ? p.bias[i] : 0.0f;
      if (v1 != nullptr) vs[p.N + i] = v1[i];
      if (p.resid_beta != nullptr) vs[2 * p.N + i] = p.resid_beta[i];
    }
  }
  const uint32_t bar_full = smem_base + L::BAR_OFF;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_tfull = bar_empty + STAGES * 8;
  const uint32_t bar_tempty = bar_tfull + 2 * 8;
  const uint32_t bar_resid = bar_tempty + 2 * 8;
  const uint32_t tmem_slot = bar_resid + 2 * EW * 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_c);
    tma_prefetch_desc(&tm_r);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + s * 8, 1);
      mbar_init(bar_tempty + s * 8, EW * CTAS);  // one arrive per epilogue warp (of both CTAs: the leader waits)
    }
    for (int s = 0; s < 2 * EW; ++s) mbar_init(bar_resid + s * 8, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CTAS == 2) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before_sync();
  if constexpr (CTAS == 2) cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // one lane, chosen with elect.sync (not `lane == 0`): TMA and tcgen05 instructions take uniform-register
    // operands, and only behind elect.sync does the compiler know a single lane is active -- otherwise it wraps
    // each of them in an ELECT / R2UR.BROADCAST / branch-back loop (see csrc/attention_bf16.cu)
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = tile_first; tile < tile_end; tile += tile_step) {
        const int m0 = static_cast<int>(tile / p.num_n_blocks) * TILE_M + static_cast<int>(rank) * BM;
        const int n0 = static_cast<int>(tile % p.num_n_blocks) * BN + static_cast<int>(rank) * (BN / CTAS);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(bar_empty + stage * 8, phase ^ 1);  // this CTA's slot is free
          if constexpr (CTAS == 2) {
            if (leader) mbar_arrive_expect_tx(bar_full + stage * 8, STAGE_TX);
            tma_load_2d_pair(&tm_a, bar_full + stage * 8, sA + stage * A_STAGE_BYTES, kb * BK, m0);
            tma_load_2d_pair(&tm_b, bar_full + stage * 8, sB + stage * L::B_STAGE_BYTES, kb * BK, n0);
          } else {
            mbar_arrive_expect_tx(bar_full + stage * 8, STAGE_TX);
            tma_load_2d(&tm_a, bar_full + stage * 8, sA + stage * A_STAGE_BYTES, kb * BK, m0);
            tma_load_2d(&tm_b, bar_full + stage * 8, sB + stage * L::B_STAGE_BYTES, kb * BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN);
      uint32_t stage = 0, phase = 0;
      uint32_t it = 0;
      for (int64_t tile = tile_first; tile < tile_end; tile += tile_step, ++it) {
        const uint32_t as = it & 1u;
        const uint32_t aphase = (it >> 1) & 1u;
        mbar_wait(bar_tempty + as * 8, aphase ^ 1);  // epilogue drained this accumulator
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(bar_full + stage * 8, phase);
          tc_fence_after_sync();
          const uint64_t a_desc = umma_desc_sw128(sA + stage * A_STAGE_BYTES);
          const uint64_t b_desc = umma_desc_sw128(sB + stage * L::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes (= 16 bf16) along K inside the swizzle atom: +2 in addr>>4 units
            if constexpr (CTAS == 2) umma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            else umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if constexpr (CTAS == 2) umma_commit_pair(bar_empty + stage * 8, 3);
          else umma_commit(bar_empty + stage * 8);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if constexpr (CTAS == 2) umma_commit_pair(bar_tfull + as * 8, 3);
        else umma_commit(bar_tfull + as * 8);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // Warp w may read TMEM lanes 32*(w%4)..+31: the EW warps split the tile into (lane quadrant,
    // column range) blocks -- 32 accumulator rows x SLABS 64-column slabs each, one row per thread.
    // bf16 outputs are staged in a private, double-buffered, 128B-swizzled [32 rows][64 cols]
    // shared-memory box and written with a TMA store, so HBM/L2 see full lines and the warps
    // never synchronise with each other.
    const int ew = warp - 4;
    const int quad = warp & 3;
    const int slab0 = (ew >> 2) * SLABS;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t my_stage = sOut + static_cast<uint32_t>(ew) * (2 * OUT_BOX_BYTES);
    const bool bf16_out = p.epilogue != HRIEMO_EPI_BIAS_RESID_F32 && p.epilogue != HRIEMO_EPI_BIAS_F32;
    // bf16 residual (pre-LayerNorm epilogue): the [32 x 64] residual box of the NEXT slab is
    // prefetched by TMA into the idle staging buffer and the sum is formed in place.
    const bool tma_resid = p.epilogue == HRIEMO_EPI_BIAS_RESID || p.epilogue == HRIEMO_EPI_BIAS_MASK;
    const bool mask_mode = p.epilogue == HRIEMO_EPI_BIAS_MASK;   // the box is the forward's post-ReLU hidden: out = box > 0 ? acc : 0
    const uint32_t my_rbar = bar_resid + static_cast<uint32_t>(ew) * 16;
    auto prefetch_resid = [&](int64_t tile_, int slab_, uint32_t ctr_) {
      const int n_ = static_cast<int>(tile_ % p.num_n_blocks) * BN + slab_ * 64;
      const int m_ = static_cast<int>((tile_ / p.num_n_blocks) * TILE_M) + static_cast<int>(rank) * BM + quad * 32;
      const uint32_t b_ = ctr_ & 1u;
      mbar_arrive_expect_tx(my_rbar + b_ * 8, OUT_BOX_BYTES);
      tma_load_2d(&tm_r, my_rbar + b_ * 8, my_stage + b_ * OUT_BOX_BYTES, n_, m_);
    };
    // first slab this warp owns in a tile (none when the tile's columns end before its range)
    auto has_slab = [&](int64_t tile_, int slab_) {
      return static_cast<int>(tile_ % p.num_n_blocks) * BN + slab_ * 64 < p.N;
    };
    // float4 of per-column vector `which` (0 bias, 1 colsum / gamma, 2 beta) at column n
    auto vec4 = [&](int which, const float* gptr, int n_) -> float4 {
      if (vec_smem) {
        float4 r;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                     : "r"(sVec + static_cast<uint32_t>(which * p.N + n_) * 4u));
        return r;
      }
      return __ldg(reinterpret_cast<const float4*>(gptr + n_));
    };
    uint32_t it = 0, slab_ctr = 0;
    const uint32_t tempty_leader = CTAS == 2 ? mapa_shared(bar_tempty, 0) : bar_tempty;
    if (tma_resid && tile_first < tile_end && has_slab(tile_first, slab0)) {
      if (elect_one()) prefetch_resid(tile_first, slab0, 0);
    }
    for (int64_t tile = tile_first; tile < tile_end; tile += tile_step, ++it) {
      const uint32_t as = it & 1u;
      const uint32_t aphase = (it >> 1) & 1u;
      const int64_t m_tile = (tile / p.num_n_blocks) * TILE_M + static_cast<int64_t>(rank) * BM;
      const int64_t m = m_tile + row_in_tile;
      const int n0 = static_cast<int>(tile % p.num_n_blocks) * BN;
      const bool row_ok = m < p.M;
      // LayerNorm folded into this GEMM: out = rstd_a * (acc - mean_a * colsum[n]) + bias[n]
      float a_mean = 0.0f, a_rstd = 1.0f;
      if (p.a_stats != nullptr && row_ok) {
        const float2 st = __ldg(p.a_stats + m);
        a_mean = st.x; a_rstd = st.y;
      }
      // residual given pre-LayerNorm: LN(r) = (r * r_scale + r_shift) * gamma[n] + beta[n]
      float2 r_st = make_float2(0.0f, 1.0f);  // consumed after the accumulator wait: the load overlaps it
      if (p.resid_stats != nullptr && row_ok) r_st = __ldg(p.resid_stats + m);
      mbar_wait(bar_tfull + as * 8, aphase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      const float r_scale = r_st.y, r_shift = -r_st.x * r_st.y;
#pragma unroll 1
      for (int sl = 0; sl < SLABS; ++sl) {
        const int s = slab0 + sl;
        const int n = n0 + s * 64;
        if (n >= p.N) break;                   // warp-uniform
        const bool two = n + 32 < p.N;         // N is a multiple of 32
        uint32_t v[2][32];
        tmem_ld32(t_row + s * 64, v[0]);
        if (two) tmem_ld32(t_row + s * 64 + 32, v[1]);
        tmem_ld_wait();
        float2 s1_2 = make_float2(0.0f, 0.0f), s2_2 = make_float2(0.0f, 0.0f);  // per-row partial LayerNorm statistics
        uint32_t buf = 0;
        if (bf16_out) {
          buf = my_stage + (slab_ctr & 1u) * OUT_BOX_BYTES;
          if (tma_resid) {
            mbar_wait(my_rbar + (slab_ctr & 1u) * 8, (slab_ctr >> 1) & 1u);  // residual box has landed
          } else {
            if (elect_one()) bulk_wait_read<1>();  // the store issued two slabs ago has left this buffer (same elected lane every time)
            __syncwarp();
          }
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (hf == 1 && !two) break;
          const int nn = n + hf * 32;
          // all elementwise math on packed fp32 pairs (FFMA2): f2[i] = columns nn + 2i, nn + 2i + 1
          float2 f2[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f2[i] = make_float2(__uint_as_float(v[hf][2 * i]), __uint_as_float(v[hf][2 * i + 1]));
          if (p.a_stats != nullptr) {
            // rstd * (acc - mean * colsum) + bias  ==  fma(rstd, acc, fma(-rstd * mean, colsum, bias))
            const float2 k1 = make_float2(a_rstd, a_rstd), k2 = make_float2(-a_rstd * a_mean, -a_rstd * a_mean);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 c4 = vec4(1, p.a_colsum, nn + i * 4);
              float4 b4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
              if (p.bias != nullptr) b4 = vec4(0, p.bias, nn + i * 4);
              f2[2 * i] = ffma2(k1, f2[2 * i], ffma2(k2, make_float2(c4.x, c4.y), make_float2(b4.x, b4.y)));
              f2[2 * i + 1] = ffma2(k1, f2[2 * i + 1], ffma2(k2, make_float2(c4.z, c4.w), make_float2(b4.z, b4.w)));
            }
          } else if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = vec4(0, p.bias, nn + i * 4);
              f2[2 * i] = fadd2(f2[2 * i], make_float2(b4.x, b4.y));
              f2[2 * i + 1] = fadd2(f2[2 * i + 1], make_float2(b4.z, b4.w));
            }
          }
          if (p.epilogue == HRIEMO_EPI_BIAS_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f2[i] = make_float2(fmaxf(f2[i].x, 0.0f), fmaxf(f2[i].y, 0.0f));
          } else if (tma_resid) {
            const float2 rs = make_float2(r_scale, r_scale), rh = make_float2(r_shift, r_shift);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t slot = static_cast<uint32_t>((hf * 4 + i) ^ (lane & 7));
              uint4 r4;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(r4.x), "=r"(r4.y), "=r"(r4.z), "=r"(r4.w)
                           : "r"(buf + lane * 128 + slot * 16)
                           : "memory");
              float2 r[4] = {make_float2(bf16_lo(r4.x), bf16_hi(r4.x)), make_float2(bf16_lo(r4.y), bf16_hi(r4.y)),
                             make_float2(bf16_lo(r4.z), bf16_hi(r4.z)), make_float2(bf16_lo(r4.w), bf16_hi(r4.w))};
              if (p.resid_stats != nullptr) {
                const float4 g0 = vec4(1, p.resid_gamma, nn + i * 8), g1 = vec4(1, p.resid_gamma, nn + i * 8 + 4);
                const float4 b0 = vec4(2, p.resid_beta, nn + i * 8), b1 = vec4(2, p.resid_beta, nn + i * 8 + 4);
                r[0] = ffma2(ffma2(r[0], rs, rh), make_float2(g0.x, g0.y), make_float2(b0.x, b0.y));
                r[1] = ffma2(ffma2(r[1], rs, rh), make_float2(g0.z, g0.w), make_float2(b0.z, b0.w));
                r[2] = ffma2(ffma2(r[2], rs, rh), make_float2(g1.x, g1.y), make_float2(b1.x, b1.y));
                r[3] = ffma2(ffma2(r[3], rs, rh), make_float2(g1.z, g1.w), make_float2(b1.z, b1.w));
              }
              if (mask_mode) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  f2[i * 4 + k] = make_float2(r[k].x > 0.0f ? f2[i * 4 + k].x : 0.0f, r[k].y > 0.0f ? f2[i * 4 + k].y : 0.0f);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) f2[i * 4 + k] = fadd2(f2[i * 4 + k], r[k]);
              }
            }
          } else if (p.epilogue == HRIEMO_EPI_BIAS_RESID_F32 && row_ok) {
            const float4* rp =
                reinterpret_cast<const float4*>(static_cast<const float*>(p.resid) + m * p.ldr + nn);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 r4 = __ldg(rp + i);
              f2[2 * i] = fadd2(f2[2 * i], make_float2(r4.x, r4.y));
              f2[2 * i + 1] = fadd2(f2[2 * i + 1], make_float2(r4.z, r4.w));
            }
          }
          if (p.stats_out != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              s1_2 = fadd2(s1_2, f2[i]);
              s2_2 = ffma2(f2[i], f2[i], s2_2);
            }
          }
          float f[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) { f[2 * i] = f2[i].x; f[2 * i + 1] = f2[i].y; }
          if (hf == 0 && tma_resid && elect_one()) {
            // Residual box of this warp's NEXT slab (same tile, or its first slab of this CTA's next
            // tile) -> the idle staging buffer.  Issued here, half a slab after the store that last
            // read that buffer, so that store has normally drained and the load has a full slab of
            // lead time instead of sitting on the critical path.
            int64_t nt_ = tile;
            int ns_ = s + 1;
            if (sl + 1 >= SLABS || n0 + ns_ * 64 >= p.N) { nt_ = tile + tile_step; ns_ = slab0; }
            if (nt_ < tile_end && has_slab(nt_, ns_)) {
              bulk_wait_read<0>();
              prefetch_resid(nt_, ns_, slab_ctr + 1);
            }
          }
          if (bf16_out) {
            // swizzled 16-byte chunks: chunk j of row r lives at slot j ^ (r & 7)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t slot = static_cast<uint32_t>((hf * 4 + g) ^ (lane & 7));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + lane * 128 + slot * 16),
                           "r"(pack_bf16(f[g * 8 + 0], f[g * 8 + 1])), "r"(pack_bf16(f[g * 8 + 2], f[g * 8 + 3])),
                           "r"(pack_bf16(f[g * 8 + 4], f[g * 8 + 5])), "r"(pack_bf16(f[g * 8 + 6], f[g * 8 + 7]))
                           : "memory");
            }
          } else if (row_ok) {  // fp32 outputs (decoder stream): direct 16-byte stores
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + m * p.ldo + nn);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              op[i] = make_float4(f[i * 4 + 0], f[i * 4 + 1], f[i * 4 + 2], f[i * 4 + 3]);
          }
        }
        if (p.stats_out != nullptr && row_ok)  // [slab][row]: consecutive lanes -> consecutive 8 bytes
          p.stats_out[static_cast<int64_t>(n >> 6) * p.M + m] = make_float2(s1_2.x + s1_2.y, s2_2.x + s2_2.y);
        if (bf16_out) {
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_store_2d(&tm_c, buf, n, static_cast<int>(m_tile) + quad * 32);
            bulk_commit();
          }
          ++slab_ctr;
        }
        __syncwarp();  // reconverge before the next warp-collective tcgen05.ld
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 2) mbar_arrive_cluster(tempty_leader + as * 8);
        else mbar_arrive(bar_tempty + as * 8);
      }
    }
    __syncwarp();
    if (elect_one()) bulk_wait_all();  // staged tiles must be out before shared memory is released
  }

  // ===================== teardown =====================
  tc_fence_before_sync();
  __syncwarp();  // role branches diverge inside warps 0 and 1; the aligned cluster barrier needs them whole
  if constexpr (CTAS == 2) cluster_sync_all();  // the peer's barriers and tensor memory stay valid until both are done
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    if constexpr (CTAS == 2) tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int BN, int STAGES, int CTAS, int EW>
static int launch_gemm(const hriemo_gemm_args& a, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES, CTAS, EW>;
  static_assert(L::DYN_BYTES <= 227 * 1024, "shared memory budget");
  CUtensorMap tm_a, tm_b, tm_c, tm_r;
  int rc = make_tmap_bf16_2d(&tm_a, a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda, BK, BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_b, a.W, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldw, BK, BN / CTAS);
  if (rc) return rc;
  const bool f32_out = a.epilogue == HRIEMO_EPI_BIAS_RESID_F32 || a.epilogue == HRIEMO_EPI_BIAS_F32;
  if (f32_out) {
    tm_c = tm_a;  // unused by the fp32 epilogues
  } else {
    rc = make_tmap_bf16_2d(&tm_c, a.out, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldo, 64, 32);
    if (rc) return rc;
  }
  tm_r = tm_a;  // only read by the bf16 residual epilogue
  if (a.epilogue == HRIEMO_EPI_BIAS_RESID || a.epilogue == HRIEMO_EPI_BIAS_MASK) {
    rc = make_tmap_bf16_2d(&tm_r, a.resid, (uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.ldr, 64, 32);
    if (rc) return rc;
  }

  GemmKernelParams p;
  p.M = a.M; p.N = a.N; p.K = a.K; p.epilogue = a.epilogue; p.bias = a.bias;
  p.out = a.out; p.ldo = a.ldo; p.resid = a.resid; p.ldr = a.ldr;
  p.a_stats = reinterpret_cast<const float2*>(a.a_stats); p.a_colsum = a.a_colsum;
  p.resid_stats = reinterpret_cast<const float2*>(a.resid_stats);
  p.resid_gamma = a.resid_gamma; p.resid_beta = a.resid_beta;
  p.stats_out = reinterpret_cast<float2*>(a.stats_out);
  p.num_n_blocks = (a.N + BN - 1) / BN;
  p.num_k_blocks = (a.K + BK - 1) / BK;
  const int64_t num_m_blocks = (a.M + BM * CTAS - 1) / (BM * CTAS);
  p.num_tiles = num_m_blocks * p.num_n_blocks;
  // HRIEMO_GEMM_WALK=contiguous | strided (default strided; read once per process: an A / B switch)
  static const int walk_env = [] { const char* e = getenv("HRIEMO_GEMM_WALK"); return e && e[0] == 'c' ? 1 : 0; }();
  p.contiguous_walk = walk_env;

  static uint64_t attr_done = 0;
  if (device_needs_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, STAGES, CTAS, EW>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess)
      return set_error(HRIEMO_ERR_CUDA, "gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const int64_t owners = device_sm_count() / CTAS;  // CTAs, or CTA pairs
  const unsigned grid = static_cast<unsigned>(p.num_tiles < owners ? p.num_tiles : owners) * CTAS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(gemm_threads(EW));
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, STAGES, CTAS, EW>, tm_a, tm_b, tm_c, tm_r, p);
  if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "gemm: launch failed: %s", cudaGetErrorString(e));
  return check_launch("gemm_bf16");
}

}  // namespace hriemo

extern "C" int hriemo_gemm_bf16(const hriemo_gemm_args* a, void* stream) {
  using namespace hriemo;
  HRIEMO_REQUIRE(a != nullptr, "gemm: null args");
  HRIEMO_REQUIRE(a->A && a->W && a->out, "gemm: null operand");
  HRIEMO_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, "gemm: bad shape M=%lld N=%d K=%d",
                 (long long)a->M, a->N, a->K);
  HRIEMO_REQUIRE(a->N % 32 == 0, "gemm: N=%d must be a multiple of 32", a->N);
  HRIEMO_REQUIRE(a->K % 8 == 0 && a->lda % 8 == 0 && a->ldw % 8 == 0,
                 "gemm: K=%d, lda=%lld, ldw=%lld must be multiples of 8", a->K, (long long)a->lda,
                 (long long)a->ldw);
  HRIEMO_REQUIRE(a->lda >= a->K && a->ldw >= a->K, "gemm: leading dimension smaller than K");
  HRIEMO_REQUIRE(a->epilogue >= HRIEMO_EPI_BIAS && a->epilogue <= HRIEMO_EPI_BIAS_MASK && a->epilogue != 4,
                 "gemm: unknown epilogue %d", a->epilogue);
  HRIEMO_REQUIRE(a->cta_pair >= 0 && a->cta_pair <= 2, "gemm: cta_pair=%d (0 auto, 1 single, 2 pair)", a->cta_pair);
  HRIEMO_REQUIRE(a->cta_pair != 2 || a->N % 256 == 0, "gemm: CTA-pair tiles need N %% 256 == 0 (N=%d)", a->N);
  const bool f32_out = a->epilogue == HRIEMO_EPI_BIAS_RESID_F32 || a->epilogue == HRIEMO_EPI_BIAS_F32;
  HRIEMO_REQUIRE(a->ldo % (f32_out ? 4 : 8) == 0, "gemm: ldo=%lld misaligned", (long long)a->ldo);
  HRIEMO_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15u) == 0, "gemm: out not 16-byte aligned");
  if (a->epilogue == HRIEMO_EPI_BIAS_RESID || a->epilogue == HRIEMO_EPI_BIAS_RESID_F32 || a->epilogue == HRIEMO_EPI_BIAS_MASK) {
    HRIEMO_REQUIRE(a->resid != nullptr, "gemm: residual epilogue without resid");
    HRIEMO_REQUIRE(a->ldr % (f32_out ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(a->resid) & 15u) == 0,
                   "gemm: resid misaligned");
  }
  if (a->bias) HRIEMO_REQUIRE((reinterpret_cast<uintptr_t>(a->bias) & 15u) == 0, "gemm: bias misaligned");
  HRIEMO_REQUIRE((a->a_stats == nullptr) == (a->a_colsum == nullptr), "gemm: a_stats and a_colsum go together");
  HRIEMO_REQUIRE(a->resid_stats == nullptr || (a->epilogue == HRIEMO_EPI_BIAS_RESID && a->resid_gamma && a->resid_beta),
                 "gemm: resid_stats needs the bf16 residual epilogue and resid_gamma / resid_beta");
  HRIEMO_REQUIRE((a->a_stats == nullptr && a->stats_out == nullptr) || !f32_out,
                 "gemm: fused LayerNorm is implemented for the bf16-output epilogues");
  HRIEMO_REQUIRE(((reinterpret_cast<uintptr_t>(a->a_stats) | reinterpret_cast<uintptr_t>(a->resid_stats) |
                   reinterpret_cast<uintptr_t>(a->stats_out)) & 7u) == 0 &&
                     ((reinterpret_cast<uintptr_t>(a->a_colsum) | reinterpret_cast<uintptr_t>(a->resid_gamma) |
                       reinterpret_cast<uintptr_t>(a->resid_beta)) & 15u) == 0,
                 "gemm: fused-LayerNorm operand misaligned");
  if (a->M == 0) return HRIEMO_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // Wide tiles for wide outputs; 128-column tiles keep more CTAs busy when N is small.
  if (a->N >= 256 && a->N % 256 == 0) {
    // CTA pairs once there is at least one 256-row tile per pair (HRIEMO_GEMM_CTAS=1|2 overrides "auto")
    static const int env_force = [] { const char* e = getenv("HRIEMO_GEMM_CTAS"); return e ? atoi(e) : 0; }();
    const int force = a->cta_pair ? a->cta_pair : env_force;
    const int64_t pair_tiles = ((a->M + 255) / 256) * (a->N / 256);
    if (force != 1 && (force == 2 || pair_tiles >= device_sm_count() / 2)) return launch_gemm<256, 5, 2, 8>(*a, s);
    return launch_gemm<256, 4, 1, 4>(*a, s);
  }
  return launch_gemm<128, 6, 1, 4>(*a, s);
}
