// Backward of hriemo_attention_bf16 on the 5th-generation tensor cores (tcgen05 + TMEM + TMA): impl = 3 of
// hriemo_attention_backward_bf16, the default where it applies.  Replaces the warp-level mma.sync passes of
// attention_bwd.cu (151 TFLOP/s, 35 % of a training step) for the backward of scaled_dot_product_attention inside the
// encoder's nn.MultiheadAttention (models/cross_modal_block_tacfn.py:74-80, 85-91, 98-104, 111-117; loss.backward() at
// scripts/fusion/train_fusion_seq_level_decoder.py:331).
//
//   P  = exp(scale * Q K^T - LSE)            (rebuilt from the forward's log-sum-exp: no running maximum, no rescale)
//   dV = P^T dO,  dP = dO V^T,  dS = scale * P o (dP - D),  D = rowsum(dO o O),  dQ = dS K,  dK = dS^T Q
//
// Two deterministic passes (no atomics) of persistent CTAs over work items (utterance, head, 128-row tile):
//   pass 0 (dK, dV): the tile holds 128 KEYS; steps of 64 queries.  S^T = K Q_step^T and dP^T = V dO_step^T accumulate
//                    side by side in tensor memory; the row-per-thread warpgroup (row = key) rebuilds P^T and dS^T with
//                    LSE / D of the step's queries as per-COLUMN vectors and stores them as bf16 over the accumulators
//                    they came from; dV += P^T dO_step and dK += dS^T Q_step then read their A operand from tensor
//                    memory and the step's rows as an MN-major shared-memory operand -- the very tile that was the
//                    K-major operand of S^T / dP^T, through a second descriptor.
//   pass 1 (dQ):     the tile holds 128 QUERIES; steps of 64 keys.  S = Q K_step^T, dP = dO V_step^T, dS as above with
//                    LSE / D per ROW, dQ += dS K_step.
// Pipeline per CTA (192 threads): warp 0 = TMA producer (four stages of step operands), warp 1 = MMA issuer (S / dP of
// step j+1 are issued into the second accumulator pair while the warpgroup works on step j), warps 2-5 = elementwise.
// Every operand is fetched through a 3-D tensor map (column, t, utterance): rows past the end of an utterance read
// as zero, so a masked probability (exactly 0) never meets a neighbour's NaN.
#include <math.h>

#include "dropout.cuh"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

constexpr int BT_M = 128;        // rows of the CTA's resident tile (UMMA M)
constexpr int BT_N = 64;         // rows of a step (UMMA N of S / dP, K extent of the accumulating MMAs)
constexpr int BT_EW = 8;         // elementwise warps: two per TMEM lane quadrant, each takes 32 of a step's 64 columns
constexpr int BT_THREADS = 64 + 32 * BT_EW;  // TMA, MMA, the elementwise warps

template <int DH>
struct BwdSmem {
  static constexpr int QCH = (DH + 63) / 64;         // 64-column (128-byte) chunks of a row
  static constexpr int R_CHUNK = BT_M * 128;         // resident tile chunk: [128 rows][128 B], 128B swizzle
  static constexpr int R_TILE = QCH * R_CHUNK;
  static constexpr int X_CHUNK = BT_N * 128;         // step operand, K-major: [64 rows][128 B]
  static constexpr int X_KM = QCH * X_CHUNK;
  // The SAME tile serves as the K-major B operand of S / dP (contraction over its columns) and as the MN-major B
  // operand of the accumulating MMAs (contraction over its rows): two descriptors over one SWIZZLE_128B tile
  // (sm100_ptx.cuh: umma_desc_sw128 / umma_desc_mn_sw128), so a step needs one copy of each operand and the ring is
  // four stages deep -- with two copies per operand only two stages fit, and the ~1.5 us TMA latency of a step's
  // operands sat on every step's critical path.
  static constexpr int STAGE = 2 * X_KM;             // x0, x1
  static constexpr int NSTAGE = 4;
  static constexpr int R_OFF = 0;                    // r0, r1
  static constexpr int X_OFF = 2 * R_TILE;
  static constexpr int BAR_OFF = X_OFF + NSTAGE * STAGE;
  static constexpr int NUM_BARS = 16;
  static constexpr int TMEM_SLOT_OFF = BAR_OFF + NUM_BARS * 8;
  static constexpr int VEC_OFF = TMEM_SLOT_OFF + 16;   // pass 0: lse2[2][64], dsum * scale [2][64] f32; pass 1: key-valid bit masks, 64 bits per step
  static int bytes(int t_k) { return VEC_OFF + 4 * 64 * 4 + ((t_k + BT_N - 1) / BT_N) * 8 + 16 + 1024; }
};

struct BwdMaps {
  CUtensorMap r0, r1;          // resident tiles (boxes of 64 rows x 64 columns, 128B swizzle)
  CUtensorMap x0, x1;          // step operands (the same boxes)
};

struct BwdParams {
  const float* lse;            // [B, H, Tq] natural log
  const float* dsum;           // [B, H, Tq]
  const uint8_t* key_pad;      // [B, Tk] or null
  const int32_t* kv_steps;     // [B] or null (pass 1 only: key steps to visit)
  __nv_bfloat16* out0;         // pass 0: dV, pass 1: dQ
  __nv_bfloat16* out1;         // pass 0: dK
  int64_t ld0, ld1;
  int B, H, Tq, Tk, n_tiles;
  float scale, scale_log2;
  uint32_t drop_p8, drop_key;   // the forward's dropout on the probabilities (0 = off): same hash, csrc/dropout.cuh
  float drop_scale;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// DROP: the instances that recompute the forward's dropout mask (a template parameter: the p = 0 step pays nothing)
template <int DH, int PASS, bool DROP = false>
__global__ void __launch_bounds__(BT_THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ BwdMaps maps, const BwdParams p) {
  using L = BwdSmem<DH>;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t ACC_COL = 256;      // [0,128) and [128,256): the two S | dP accumulator pairs; accumulators behind
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t sR0 = base + L::R_OFF, sR1 = sR0 + L::R_TILE, sX = base + L::X_OFF;
  const uint32_t bars = base + L::BAR_OFF;
  const uint32_t b_rfull = bars;              // resident tiles landed
  const uint32_t b_xfull = bars + 1 * 8;      // [stage 4]
  const uint32_t b_xempty = bars + 5 * 8;     // [stage 4]
  const uint32_t b_sfull = bars + 9 * 8;      // [accumulator pair]: S and dP of a step are complete
  const uint32_t b_edone = bars + 11 * 8;     // [accumulator pair]: P / dS of a step are in tensor memory
  const uint32_t b_accfull = bars + 13 * 8;   // the tile's accumulators are complete
  const uint32_t b_rempty = bars + 14 * 8;    // the tile's S / dP MMAs have read the resident tiles
  const uint32_t b_accempty = bars + 15 * 8;  // the epilogue has read the accumulators
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::TMEM_SLOT_OFF);
  float* vec = reinterpret_cast<float*>(base_ptr + L::VEC_OFF);
  uint8_t* kvalid = base_ptr + L::VEC_OFF + 4 * 64 * 4;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Persistent: a CTA walks work items (utterance, head, 128-row tile); tensor memory, barriers and the pipeline state
  // (global step counter g -> ring stage, accumulator pair and their phases; item counter -> resident / accumulator
  // phases) live across items, so an item costs its steps and not a CTA launch -- the cross-attention shapes have
  // ONE step per item.
  const uint32_t n_items = static_cast<uint32_t>(p.n_tiles) * p.H * p.B;
  const int t_step = PASS == 0 ? p.Tq : p.Tk;         // the sequence the steps walk
  auto decode = [&](uint32_t it, int& tile, int& h, int& b) {
    tile = static_cast<int>(it % static_cast<uint32_t>(p.n_tiles));
    const uint32_t r = it / static_cast<uint32_t>(p.n_tiles);
    h = static_cast<int>(r % static_cast<uint32_t>(p.H));
    b = static_cast<int>(r / static_cast<uint32_t>(p.H));
  };
  auto steps_of = [&](int b) {
    int n = (t_step + BT_N - 1) / BT_N;
    if (PASS == 1 && p.kv_steps != nullptr) n = min(n, __ldg(p.kv_steps + b));
    return n;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.r0);
    tma_prefetch_desc(&maps.r1);
    tma_prefetch_desc(&maps.x0);
    tma_prefetch_desc(&maps.x1);
    mbar_init(b_rfull, 1);
    mbar_init(b_rempty, 1);
    for (int s = 0; s < L::NSTAGE; ++s) {
      mbar_init(b_xfull + s * 8, 1);
      mbar_init(b_xempty + s * 8, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_sfull + s * 8, 1);
      mbar_init(b_edone + s * 8, BT_EW);   // one arrival per elementwise warp
    }
    mbar_init(b_accfull, 1);
    mbar_init(b_accempty, BT_EW);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(base + L::TMEM_SLOT_OFF);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t g = 0, ri = 0;
      for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x, ++ri) {
        int tile, h, b;
        decode(it, tile, h, b);
        const int row0 = tile * BT_M;
        const int n_steps = steps_of(b);
        mbar_wait(b_rempty, (ri & 1u) ^ 1u);     // the previous item's S / dP MMAs are done with the resident tiles
        mbar_arrive_expect_tx(b_rfull, 2 * L::R_TILE);
        for (int c = 0; c < L::QCH; ++c)
          for (int half = 0; half < 2; ++half) {
            tma_load_3d(&maps.r0, b_rfull, sR0 + c * L::R_CHUNK + half * L::X_CHUNK, h * DH + c * 64, row0 + half * BT_N, b);
            tma_load_3d(&maps.r1, b_rfull, sR1 + c * L::R_CHUNK + half * L::X_CHUNK, h * DH + c * 64, row0 + half * BT_N, b);
          }
        for (int j = 0; j < n_steps; ++j, ++g) {
          const uint32_t s = g % L::NSTAGE, par = (g / L::NSTAGE) & 1u;
          const uint32_t st = sX + s * L::STAGE;
          mbar_wait(b_xempty + s * 8, par ^ 1);
          mbar_arrive_expect_tx(b_xfull + s * 8, L::STAGE);
          for (int c = 0; c < L::QCH; ++c) {
            tma_load_3d(&maps.x0, b_xfull + s * 8, st + c * L::X_CHUNK, h * DH + c * 64, j * BT_N, b);
            tma_load_3d(&maps.x1, b_xfull + s * 8, st + L::X_KM + c * L::X_CHUNK, h * DH + c * 64, j * BT_N, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      if (tmem_base != 0u) __trap();   // all 512 columns are this CTA's: the allocation starts at 0 (uniform operands, see attention_bf16.cu)
      constexpr uint32_t idesc_s = umma_idesc_bf16(BT_M, BT_N);
      constexpr uint32_t idesc_acc = umma_idesc_bf16(BT_M, DH) | kUmmaBMajorMN;
      const uint64_t r0_desc = umma_desc_sw128(sR0), r1_desc = umma_desc_sw128(sR1);
      // S | dP of global step g into accumulator pair g & 1
      auto issue_sdp = [&](uint32_t g) {
        const uint32_t xs = g % L::NSTAGE, xpar = (g / L::NSTAGE) & 1u;
        const uint32_t s = g & 1u;
        const uint32_t st = sX + xs * L::STAGE;
        mbar_wait(b_xfull + xs * 8, xpar);
        tc_fence_after_sync();
        const uint64_t x0_desc = umma_desc_sw128(st), x1_desc = umma_desc_sw128(st + L::X_KM);
        const uint32_t d_s = s * 128u, d_dp = s * 128u + 64u;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16(d_s, r0_desc + (((k >> 2) * L::R_CHUNK + (k & 3) * 32) >> 4),
                    x0_desc + (((k >> 2) * L::X_CHUNK + (k & 3) * 32) >> 4), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16(d_dp, r1_desc + (((k >> 2) * L::R_CHUNK + (k & 3) * 32) >> 4),
                    x1_desc + (((k >> 2) * L::X_CHUNK + (k & 3) * 32) >> 4), idesc_s, k != 0);
        umma_commit(b_sfull + s * 8);
      };
      uint32_t g = 0, ri = 0;
      bool pre_issued = false;   // the first S | dP of this item went out during the previous item's last step
      for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x, ++ri) {
        int tile, h, b;
        decode(it, tile, h, b);
        const int n_steps = steps_of(b);
        if (!pre_issued) {
          mbar_wait(b_rfull, ri & 1u);
          issue_sdp(g);
          if (n_steps == 1) umma_commit(b_rempty);
        }
        pre_issued = false;
        for (int j = 0; j < n_steps; ++j, ++g) {
          const uint32_t s = g & 1u, par = (g >> 1) & 1u;
          const uint32_t xs = g % L::NSTAGE;
          const uint32_t st = sX + xs * L::STAGE;
          if (j + 1 < n_steps) {
            issue_sdp(g + 1);   // the other accumulator pair: its last readers (step g-1) were issued before
            if (j + 2 == n_steps) umma_commit(b_rempty);   // the item's last S / dP: the resident tiles may be replaced
          } else if (it + gridDim.x < n_items) {
            // (v22) last step of the item: look ahead ACROSS the item boundary.  The resident tiles were handed back with
            // this item's last S | dP, so the producer is already loading the next item's; its first S | dP go into the
            // other accumulator pair while the elementwise warps are on this step -- the cross-attention shapes have ONE
            // step per item and ran [S | dP -> elementwise -> accumulate -> drain] strictly one after the other.
            int tile2, h2, b2;
            decode(it + gridDim.x, tile2, h2, b2);
            mbar_wait(b_rfull, (ri + 1u) & 1u);
            issue_sdp(g + 1);
            if (steps_of(b2) == 1) umma_commit(b_rempty);
            pre_issued = true;
          }
          mbar_wait(b_edone + s * 8, par);
          if (j == 0) mbar_wait(b_accempty, (ri & 1u) ^ 1u);   // the previous item's epilogue has read the accumulators
          tc_fence_after_sync();
          // the step's rows again, now as MN-major operands (contraction over the 64 rows, 16 rows = 2 048 B per MMA)
          const uint64_t x0_mn = umma_desc_mn_sw128(st, L::X_CHUNK);
          if (PASS == 0) {
            const uint64_t x1_mn = umma_desc_mn_sw128(st + L::X_KM, L::X_CHUNK);
#pragma unroll
            for (int k = 0; k < BT_N / 16; ++k)    // dV += P^T dO_step
              umma_bf16_ts(ACC_COL, s * 128u + k * 8 + (k >= 2 ? 16 : 0), x1_mn + ((k * 16 * 128) >> 4), idesc_acc, (j | k) != 0);
#pragma unroll
            for (int k = 0; k < BT_N / 16; ++k)    // dK += dS^T Q_step
              umma_bf16_ts(ACC_COL + DH, s * 128u + 64u + k * 8 + (k >= 2 ? 16 : 0), x0_mn + ((k * 16 * 128) >> 4), idesc_acc, (j | k) != 0);
          } else {
#pragma unroll
            for (int k = 0; k < BT_N / 16; ++k)    // dQ += dS K_step
              umma_bf16_ts(ACC_COL, s * 128u + 64u + k * 8 + (k >= 2 ? 16 : 0), x0_mn + ((k * 16 * 128) >> 4), idesc_acc, (j | k) != 0);
          }
          umma_commit(b_xempty + xs * 8);
        }
        umma_commit(b_accfull);
      }
    }
  } else {
    // ===================== elementwise warps: one row of the tile per thread, TWO threads per row =====================
    // (v21) warps 2-5 take columns [0, 32) of every step, warps 6-9 columns [32, 64): with one warp per scheduler a step's
    // 64 exponentials + dS per thread (~1 800 cycles with the tcgen05.ld / st round trips) were twice the step's ~980
    // cycles of tensor work.  No exchange is needed between the two (the forward's LSE makes P elementwise); each warp
    // stores its bf16 P / dS over the first 16 columns of the 32 IT read (columns [0,16) and [32,48) of the accumulator:
    // the MMA issuer addresses the two K = 32 halves of the A operand separately).
    const int half = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int wt = (warp - 2) * 32 + lane;          // thread within the warpgroup
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float sc2 = p.scale_log2;
    constexpr float LOG2E = 1.4426950408889634f;
    uint64_t* kmask = reinterpret_cast<uint64_t*>(kvalid);   // pass 1: bit c of kmask[j] = key j * 64 + c takes part
    const float2 sc2v = make_float2(sc2, sc2), scv = make_float2(p.scale, p.scale);
    uint32_t g = 0, ri = 0;
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x, ++ri) {
      int tile, h, b;
      decode(it, tile, h, b);
      const int row0 = tile * BT_M;
      const int n_steps = steps_of(b);
      const float* lse_bh = p.lse + (static_cast<int64_t>(b) * p.H + h) * p.Tq;
      const float* dsum_bh = p.dsum + (static_cast<int64_t>(b) * p.H + h) * p.Tq;
      bool row_valid;           // pass 0: this thread's key takes part; pass 1: this thread's query exists
      float lse2_row = 0.0f, dsc_row = 0.0f;
      if (PASS == 0) {
        const int k = row0 + row;
        row_valid = k < p.Tk && (p.key_pad == nullptr || p.key_pad[static_cast<int64_t>(b) * p.Tk + k] == 0);
      } else {
        const int q = row0 + row;
        row_valid = q < p.Tq;
        lse2_row = row_valid ? __ldg(lse_bh + q) * LOG2E : INFINITY;   // +inf => P = 0
        dsc_row = row_valid ? __ldg(dsum_bh + q) * p.scale : 0.0f;
        // (every warp is past the previous item's last step here: its accumulators only completed after all four
        // warps had handed that step over)
        for (int j = wt; j < n_steps; j += 32 * BT_EW) {
          uint64_t m = 0;
          for (int c = 0; c < BT_N; ++c) {
            const int k = j * BT_N + c;
            if (k < p.Tk && (p.key_pad == nullptr || p.key_pad[static_cast<int64_t>(b) * p.Tk + k] == 0)) m |= 1ull << c;
          }
          kmask[j] = m;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * BT_EW) : "memory");
      }
      const float rowmask = row_valid ? 1.0f : 0.0f;
      const float2 rmv = make_float2(rowmask, rowmask);
      // dropout on the probabilities: pass 0 owns key (row0 + row) and walks queries (one hash per query: word k >> 2,
      // byte k & 3), pass 1 owns query (row0 + row) and walks the words of its row
      [[maybe_unused]] uint32_t dfix = 0u, dshift = 0u;
      if constexpr (DROP) {
        const uint32_t dkey = drop_key_bh(p.drop_key, static_cast<uint32_t>(b * p.H + h));
        dfix = PASS == 0 ? dkey + (static_cast<uint32_t>(row0 + row) >> 2) * DROP_C_WORD
                         : dkey + static_cast<uint32_t>(row0 + row) * DROP_C_ROW;
        dshift = (static_cast<uint32_t>(row0 + row) & 3u) * 8u;   // pass 0: this key's byte
      }
      // pass 0: LSE / D of a step's 64 queries as per-column vectors in shared memory (lse * log2 e, +inf past Tq so
      // that P = 0 there; D * scale), double-buffered; the NEXT step's values are fetched from global memory while
      // this step computes and only touched (scaled, stored) a whole step later
      float pre = 0.0f;
      const float* vsrc = wt < 64 ? lse_bh : dsum_bh;
      const float vmul = wt < 64 ? LOG2E : p.scale;
      auto fetch_vec = [&](int j) {
        if (wt >= 128) return;   // the first four warps fetch / stash; all eight read
        const int q = j * BT_N + (wt & 63);
        pre = q < p.Tq ? __ldg(vsrc + q) : (wt < 64 ? INFINITY : 0.0f);
      };
      auto stash_vec = [&](uint32_t gg) { if (wt < 128) vec[(wt < 64 ? 0 : 128) + (gg & 1u) * 64 + (wt & 63)] = pre * vmul; };
      if (PASS == 0) {
        fetch_vec(0);
        stash_vec(g);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * BT_EW) : "memory");
      }
      for (int j = 0; j < n_steps; ++j, ++g) {
        const uint32_t s = g & 1u, par = (g >> 1) & 1u;
        const float4* lse2_c = reinterpret_cast<const float4*>(vec + s * 64);          // pass 0: this step's column vectors
        const float4* dsc_c = reinterpret_cast<const float4*>(vec + 128 + s * 64);
        if (PASS == 0 && j + 1 < n_steps) fetch_vec(j + 1);
        const uint64_t km = PASS == 1 ? kmask[j] : ~0ull;
        mbar_wait(b_sfull + s * 8, par);
        tc_fence_after_sync();
        const uint32_t t_s = t_row + s * 128u, t_dp = t_s + 64u;
        {
          uint32_t vs[32], vd[32];
          tmem_ld32(t_s + half * 32, vs);
          tmem_ld32(t_dp + half * 32, vd);
          tmem_ld_wait();
          uint32_t pp[16], pd[16];
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {         // four columns at a time
            float2 l01, l23, d01, d23;
            if (PASS == 0) {
              const float4 l = lse2_c[half * 8 + i4], dd = dsc_c[half * 8 + i4];
              l01 = make_float2(-l.x, -l.y); l23 = make_float2(-l.z, -l.w);
              d01 = make_float2(-dd.x, -dd.y); d23 = make_float2(-dd.z, -dd.w);
            } else {
              l01 = l23 = make_float2(-lse2_row, -lse2_row);
              d01 = d23 = make_float2(-dsc_row, -dsc_row);
            }
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
              const int i = i4 * 2 + pr;           // column pair (2 i, 2 i + 1) of this half
              const float2 x = ffma2(make_float2(__uint_as_float(vs[2 * i]), __uint_as_float(vs[2 * i + 1])), sc2v, pr ? l23 : l01);
              float2 pv = make_float2(ex2f(x.x), ex2f(x.y));
              if (PASS == 0) {
                pv = fmul2(pv, rmv);               // a PAD key's row is switched off as a whole
              } else if (km != ~0ull) {            // (warp-uniform) some key of this step is PAD or past Tk
                const int c = half * 32 + 2 * i;
                if (!((km >> c) & 1ull)) pv.x = 0.0f;
                if (!((km >> (c + 1)) & 1ull)) pv.y = 0.0f;
              }
              if constexpr (DROP) {   // O = (P o M / (1 - p)) V: dV takes P o M c, dP = (dO V^T) o M c; dS keeps the undropped P
                bool k0, k1;
                const int c = j * BT_N + half * 32 + 2 * i;     // column of the pair: a query (pass 0) or a key (pass 1)
                if (PASS == 0) {
                  k0 = ((drop_mix(dfix + static_cast<uint32_t>(c) * DROP_C_ROW) >> dshift) & 0xffu) >= p.drop_p8;
                  k1 = ((drop_mix(dfix + static_cast<uint32_t>(c + 1) * DROP_C_ROW) >> dshift) & 0xffu) >= p.drop_p8;
                } else {
                  const uint32_t word = drop_mix(dfix + (static_cast<uint32_t>(c) >> 2) * DROP_C_WORD);
                  k0 = ((word >> ((c & 3) * 8)) & 0xffu) >= p.drop_p8;
                  k1 = ((word >> ((c & 3) * 8 + 8)) & 0xffu) >= p.drop_p8;
                }
                const float2 mc = make_float2(k0 ? p.drop_scale : 0.0f, k1 ? p.drop_scale : 0.0f);
                const float2 pvm = fmul2(pv, mc);   // what multiplied V in the forward
                const float2 dp = fmul2(make_float2(__uint_as_float(vd[2 * i]), __uint_as_float(vd[2 * i + 1])), mc);
                const float2 t = ffma2(dp, scv, pr ? d23 : d01);
                const float2 ds = fmul2(pv, t);
                pp[i] = pack_bf16(pvm.x, pvm.y);
                pd[i] = pack_bf16(ds.x, ds.y);
              } else {
                // dS = scale * P o (dP - D) = P o (dP * scale - D * scale)
                const float2 t = ffma2(make_float2(__uint_as_float(vd[2 * i]), __uint_as_float(vd[2 * i + 1])), scv, pr ? d23 : d01);
                const float2 ds = fmul2(pv, t);
                pp[i] = pack_bf16(pv.x, pv.y);
                pd[i] = pack_bf16(ds.x, ds.y);
              }
            }
          }
          // bf16 P over the first 16 of the 32 S columns this warp read, bf16 dS likewise over dP: in registers before the store
          if (PASS == 0) tmem_st16(t_s + half * 32, pp);
          tmem_st16(t_dp + half * 32, pd);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_edone + s * 8);
        if (PASS == 0 && j + 1 < n_steps) {
          stash_vec(g + 1);                                        // (that buffer was last read in step g-1)
          asm volatile("bar.sync 1, %0;" ::"n"(32 * BT_EW) : "memory");
        }
      }
      // ---- epilogue: accumulators -> registers (then the MMA issuer may start the next item) -> bf16 -> global
      mbar_wait(b_accfull, ri & 1u);
      tc_fence_after_sync();
      const int t_out = PASS == 0 ? p.Tk : p.Tq;
      const bool store = row0 + row < t_out;
      // the two warps of a lane quadrant share the accumulators' columns: 32-column chunks dealt alternately
      constexpr int N_ACC = PASS == 0 ? 2 : 1;
      constexpr int N_CHUNK = N_ACC * (DH / 32);
      bool arrived = false;
#pragma unroll
      for (int ch = 0; ch < N_CHUNK; ++ch) {
        if ((ch & 1) != half) continue;
        const int a = ch / (DH / 32), c0 = (ch % (DH / 32)) * 32;
        __nv_bfloat16* dst = (a == 0 ? p.out0 : p.out1) +
                             (static_cast<int64_t>(b) * t_out + row0 + row) * (a == 0 ? p.ld0 : p.ld1) + h * DH;
        uint32_t v[32];
        tmem_ld32(t_row + ACC_COL + a * DH + c0, v);
        tmem_ld_wait();
        if (ch + 2 >= N_CHUNK) {     // this warp's last accumulator columns are in registers
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(b_accempty);
          arrived = true;
        }
        if (store) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v[g4 * 8 + 0]), __uint_as_float(v[g4 * 8 + 1]));
            o.y = pack_bf16(__uint_as_float(v[g4 * 8 + 2]), __uint_as_float(v[g4 * 8 + 3]));
            o.z = pack_bf16(__uint_as_float(v[g4 * 8 + 4]), __uint_as_float(v[g4 * 8 + 5]));
            o.w = pack_bf16(__uint_as_float(v[g4 * 8 + 6]), __uint_as_float(v[g4 * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + c0 + g4 * 8) = o;
          }
        }
      }
      if (!arrived) {                // (a single chunk, dh = 32 in pass 1: the second warp has nothing to read)
        __syncwarp();
        if (lane == 0) mbar_arrive(b_accempty);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

static int make_km(CUtensorMap* m, const void* base, int d, int T, int B, int64_t ld) {
  const uint64_t dims[3] = {(uint64_t)d, (uint64_t)T, (uint64_t)B};
  const uint64_t pitch[2] = {(uint64_t)ld, (uint64_t)ld * T};
  const uint32_t box[3] = {64u, (uint32_t)BT_N, 1u};
  return make_tmap_bf16_3d_plain(m, base, dims, pitch, box, 128);
}

template <int DH>
int launch_attn_bwd_tc(const hriemo_attn_bwd_args& a, cudaStream_t stream) {
  using L = BwdSmem<DH>;
  const int d = a.H * DH;
  const int smem = L::bytes(a.Tk);
  if (smem > 227 * 1024)
    return set_error(HRIEMO_ERR_INVALID, "attention_backward (tcgen05): Tk=%d too long for the key-valid bytes (dh=%d)", a.Tk, DH);
  BwdParams p;
  p.lse = a.lse; p.dsum = a.dsum; p.key_pad = a.key_pad; p.kv_steps = a.kv_steps;
  p.B = a.B; p.H = a.H; p.Tq = a.Tq; p.Tk = a.Tk;
  p.scale = a.scale; p.scale_log2 = a.scale * 1.4426950408889634f;
  p.drop_p8 = a.drop_p8; p.drop_key = a.drop_key; p.drop_scale = a.drop_p8 ? a.drop_scale : 1.0f;
  static uint64_t attr_done = 0;
  if (device_needs_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_tc_kernel<DH, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_tc_kernel<DH, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_tc_kernel<DH, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_tc_kernel<DH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess)
      return set_error(HRIEMO_ERR_CUDA, "attention_backward (tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  int rc;
  {
    // pass 0: dK / dV.  resident = K, V; steps = Q, dO
    BwdMaps m;
    if ((rc = make_km(&m.r0, a.k, d, a.Tk, a.B, a.ldk))) return rc;
    if ((rc = make_km(&m.r1, a.v, d, a.Tk, a.B, a.ldv))) return rc;
    if ((rc = make_km(&m.x0, a.q, d, a.Tq, a.B, a.ldq))) return rc;
    if ((rc = make_km(&m.x1, a.d_out, d, a.Tq, a.B, a.lddo))) return rc;
    p.out0 = static_cast<__nv_bfloat16*>(a.dv); p.ld0 = a.lddv;
    p.out1 = static_cast<__nv_bfloat16*>(a.dk); p.ld1 = a.lddk;
    p.n_tiles = (a.Tk + BT_M - 1) / BT_M;
    const int64_t items = static_cast<int64_t>(p.n_tiles) * a.H * a.B;
    if (items >= (1ll << 31)) return set_error(HRIEMO_ERR_INVALID, "attention_backward (tcgen05): too many work items");
    const unsigned grid = static_cast<unsigned>(items < device_sm_count() ? items : device_sm_count());
    if (p.drop_p8) attention_bwd_tc_kernel<DH, 0, true><<<grid, BT_THREADS, smem, stream>>>(m, p);
    else attention_bwd_tc_kernel<DH, 0><<<grid, BT_THREADS, smem, stream>>>(m, p);
    if ((rc = check_launch("attention_backward (tcgen05, dK / dV)"))) return rc;
  }
  {
    // pass 1: dQ.  resident = Q, dO; steps = K, V
    BwdMaps m;
    if ((rc = make_km(&m.r0, a.q, d, a.Tq, a.B, a.ldq))) return rc;
    if ((rc = make_km(&m.r1, a.d_out, d, a.Tq, a.B, a.lddo))) return rc;
    if ((rc = make_km(&m.x0, a.k, d, a.Tk, a.B, a.ldk))) return rc;
    if ((rc = make_km(&m.x1, a.v, d, a.Tk, a.B, a.ldv))) return rc;
    p.out0 = static_cast<__nv_bfloat16*>(a.dq); p.ld0 = a.lddq;
    p.out1 = nullptr; p.ld1 = 0;
    p.n_tiles = (a.Tq + BT_M - 1) / BT_M;
    const int64_t items = static_cast<int64_t>(p.n_tiles) * a.H * a.B;
    if (items >= (1ll << 31)) return set_error(HRIEMO_ERR_INVALID, "attention_backward (tcgen05): too many work items");
    const unsigned grid = static_cast<unsigned>(items < device_sm_count() ? items : device_sm_count());
    if (p.drop_p8) attention_bwd_tc_kernel<DH, 1, true><<<grid, BT_THREADS, smem, stream>>>(m, p);
    else attention_bwd_tc_kernel<DH, 1><<<grid, BT_THREADS, smem, stream>>>(m, p);
    if ((rc = check_launch("attention_backward (tcgen05, dQ)"))) return rc;
  }
  return HRIEMO_OK;
}

template int launch_attn_bwd_tc<32>(const hriemo_attn_bwd_args&, cudaStream_t);
template int launch_attn_bwd_tc<64>(const hriemo_attn_bwd_args&, cudaStream_t);
template int launch_attn_bwd_tc<96>(const hriemo_attn_bwd_args&, cudaStream_t);
template int launch_attn_bwd_tc<128>(const hriemo_attn_bwd_args&, cudaStream_t);

}  // namespace hriemo
