// Flash-style multi-head attention forward for sm_100a (tcgen05 + TMEM + TMA), v3:
//   O[b, tq, h*dh:(h+1)*dh] = softmax_k(Q K^T * scale + key_padding) V
//
// Persistent kernel, one CTA per SM, walking a flat sequence of steps
// (work item = (utterance, head, pair of 128-row query tiles)) x (64-key step).  352 threads:
//   warp 0       TMA producer: Q tiles (double-buffered per query tile), K / V tiles of 64 keys
//                in a ring; it runs ahead across work items.
//   warps 1, 10  tcgen05.mma issuers, one per query tile, so the two tiles never wait on each other.
//                S_t = Q_t K^T is DOUBLE-BUFFERED in tensor memory and issued two steps ahead of
//                the softmax (also across work-item boundaries); O_t += P_t V reads P_t from tensor
//                memory (P aliases the S buffer it was computed from) and V from shared memory as an
//                MN-major operand (row-major [key][column] tiles exactly as the projection GEMM wrote
//                them: no transposed copy of V exists anywhere).
//   warps 2..5   softmax warpgroup of query tile 0 \ one query row per thread (TMEM lane = row):
//   warps 6..9   softmax warpgroup of query tile 1 / 64 scores held in registers, row max, lazy
//                running max (O / l rescaled only when it grows by more than 2^8),
//                p = ex2(fma(s, scale*log2e, -m)), P stored as bf16 into TMEM, fp32 row sums;
//                final O / l written as bf16.
// Tensor memory per query tile: S0 [0,64) S1 [64,128) O [128,128+dh)  -> 2 x 256 columns.
//
// Short query sequences (Tq <= 128, H even: text self-attention, text->audio, MOSEI text) have ONE query
// tile per (utterance, head); in the form above the second tile's MMA issuer and softmax warpgroup would
// idle.  In PAIRED-HEAD mode a work item is (utterance, pair of heads): tile t carries head 2h'+t, and the
// item's flat steps alternate between the two heads (flat step f: tile f & 1, key step f >> 1), so the two
// tiles share the K / V ring in time instead of sharing its contents.  Each role skips the flat steps of the
// other tile exactly as it skips a tile that does not exist; S buffers are indexed by the tile's own step count.
//
// Replaces the scaled_dot_product_attention inside nn.MultiheadAttention at
// models/cross_modal_block_tacfn.py:74-80,85-91,98-104,111-117 and
// models/cross_modal_block.py:56-59,64-67 of the reference.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "dropout.cuh"
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

constexpr int A3_BQ = 128;   // query rows per tile (UMMA M)
constexpr int A3_BKV = 64;   // keys per step (UMMA N of S, K extent of PV)
constexpr int A3_THREADS = 352;  // TMA, MMA(tile 0), 2 x 4 softmax warps, MMA(tile 1)
constexpr float A3_LAZY_TAU = 8.0f;  // log2 units

template <int DH>
struct Attn3Smem {
  static constexpr int QCH = (DH + 63) / 64;       // 64-column (128-byte) chunks of Q / K rows
  static constexpr int Q_CHUNK = A3_BQ * 128;      // [128 rows][128 B], 128B-swizzled
  static constexpr int Q_TILE = QCH * Q_CHUNK;
  static constexpr int K_CHUNK = A3_BKV * 128;     // [64 keys][128 B]
  static constexpr int K_STAGE = QCH * K_CHUNK;
  static constexpr int V_GROUP = A3_BKV * 64;      // [64 keys][32 columns = 64 B], 64B-swizzled
  static constexpr int V_STAGE = (DH / 32) * V_GROUP;
  static constexpr int KV_STAGES = (DH > 96) ? 2 : 3;
  static constexpr int Q_OFF = 0;                  // [tile 2][buffer 2]
  static constexpr int K_OFF = Q_OFF + 4 * Q_TILE;
  static constexpr int V_OFF = K_OFF + KV_STAGES * K_STAGE;
  static constexpr int BAR_OFF = V_OFF + KV_STAGES * V_STAGE;
  // q_full[2][2] q_empty[2][2] k_full[3] k_empty[3] v_full[3] v_empty[3] s_full[2][2] p_full[2][2] pv_done[2][2] s_free[2][2]
  static constexpr int NUM_BARS = 36;
  static constexpr int TMEM_SLOT_OFF = BAR_OFF + NUM_BARS * 8;
  static constexpr int DYN_OFF = TMEM_SLOT_OFF + 16;   // caps[2][n_kv*64] f32, flags[2][n_kv] i32
  static int dyn_bytes(int n_kv) { return DYN_OFF + 2 * n_kv * A3_BKV * 4 + 2 * n_kv * 4 + 1024; }
};

// Pipeline trace (tools/attn_trace.py builds with -DHRIEMO_ATTN_TRACE): CTA 0 records clock64() at fixed
// points of its first 64 steps, per role (MMA issuer 0/1, softmax warpgroup 0/1).  Compiled out otherwise.
#ifdef HRIEMO_ATTN_TRACE
__device__ long long* g_attn_trace = nullptr;
#define ATRACE(role, step, ev)                                                                       \
  do {                                                                                               \
    if (blockIdx.x == 0 && g_attn_trace != nullptr && (step) < 64u)                                  \
      g_attn_trace[((role) * 64 + (step)) * 8 + (ev)] = clock64();                                    \
  } while (0)
#else
#define ATRACE(role, step, ev) do {} while (0)
#endif

struct Attn3Params {
  const int32_t* kv_steps;   // [B] key steps to run per utterance (trailing all-PAD tiles skipped), or null
  const uint8_t* key_pad;
  __nv_bfloat16* out;
  float* lse;                // [B, H, Tq] natural-log sum of exponentials of the scaled scores, or null
  int64_t ldo;
  int B, H, Tq, Tk;
  int n_kv, n_qp;
  int paired;       // paired-head mode (see the header): items are (utterance, head pair)
  int items_per_b;  // work items per utterance: H * n_qp, or H / 2 in paired-head mode
  int contiguous;   // item -> CTA assignment, see the kernel
  int64_t n_items;
  float scale_log2;
  uint32_t drop_p8, drop_key;   // dropout on the probabilities (training): 0 = off
  float drop_scale;             // 1 / (1 - p), folded into the final 1 / l
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// score = fminf(s, cap): cap is +inf for a valid key and -inf for a masked one; fminf returns the
// non-NaN operand, so garbage (even NaN) in a masked column becomes exactly -inf.
__device__ __forceinline__ void apply_caps(uint32_t (&v)[32], const float* cap) {
  const float4* cp = reinterpret_cast<const float4*>(cap);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 ck = cp[i];
    v[i * 4 + 0] = __float_as_uint(fminf(__uint_as_float(v[i * 4 + 0]), ck.x));
    v[i * 4 + 1] = __float_as_uint(fminf(__uint_as_float(v[i * 4 + 1]), ck.y));
    v[i * 4 + 2] = __float_as_uint(fminf(__uint_as_float(v[i * 4 + 2]), ck.z));
    v[i * 4 + 3] = __float_as_uint(fminf(__uint_as_float(v[i * 4 + 3]), ck.w));
  }
}

__device__ __forceinline__ void max4(float (&m4)[4], const uint32_t (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    m4[0] = fmaxf(m4[0], __uint_as_float(v[i + 0]));
    m4[1] = fmaxf(m4[1], __uint_as_float(v[i + 1]));
    m4[2] = fmaxf(m4[2], __uint_as_float(v[i + 2]));
    m4[3] = fmaxf(m4[3], __uint_as_float(v[i + 3]));
  }
}

// p = 2^(s*sc + neg_m) for 32 scores -> 16 packed bf16 pairs; the scale/shift and the row sums run on
// packed fp32 pairs (FFMA2 / FADD2: half the instructions of the scalar form), two sum chains.
__device__ __forceinline__ void exp_pack(const uint32_t (&v)[32], float sc, float neg_m, float (&l4)[4],
                                         uint32_t (&pk)[16]) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(neg_m, neg_m);
  float2 la = make_float2(l4[0], l4[2]), lb = make_float2(l4[1], l4[3]);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = ffma2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
    const float2 e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
    pk[i] = pack_bf16(e.x, e.y);
    if (i & 1) lb = fadd2(lb, e); else la = fadd2(la, e);
  }
  l4[0] = la.x; l4[2] = la.y; l4[1] = lb.x; l4[3] = lb.y;
}


// The two halves of exp_pack for the v5 softmax loop: x = s*sc + neg_m for 32 scores (after which the raw scores are
// dead and their registers can take the next step's tensor-memory load), and p = 2^x -> 16 packed bf16 pairs + row sums.
__device__ __forceinline__ void scale32(const uint32_t (&v)[32], float sc, float neg_m, float2 (&x)[16]) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(neg_m, neg_m);
#pragma unroll
  for (int i = 0; i < 16; ++i)
    x[i] = ffma2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
}
// (volatile: the exponentials stay between the turn-taking barriers and behind the tensor-memory loads they are to cover;
// without it the compiler hoists the first chunk's exponentials above the barrier -- HRIEMO_ATTN_V5_LOOSE keeps that form
// for A/B runs)
__device__ __forceinline__ float ex2_approx_v(float x) {
  float y;
#ifdef HRIEMO_ATTN_V5_LOOSE
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
#else
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
#endif
  return y;
}
__device__ __forceinline__ void exp_pack_x(const float2 (&x)[16], float (&l4)[4], uint32_t (&pk)[16]) {
  float2 la = make_float2(l4[0], l4[2]), lb = make_float2(l4[1], l4[3]);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 e = make_float2(ex2_approx_v(x[i].x), ex2_approx_v(x[i].y));
    pk[i] = pack_bf16(e.x, e.y);
    if (i & 1) lb = fadd2(lb, e); else la = fadd2(la, e);
  }
  l4[0] = la.x; l4[2] = la.y; l4[1] = lb.x; l4[3] = lb.y;
}

// Dropout on 32 probabilities packed as bf16 pairs (pk[i] = keys 2 i, 2 i + 1 of the chunk whose first hash word is w0):
// dropped ones become exact zeros; the row sums were taken before (torch normalises first and drops afterwards).
__device__ __forceinline__ void drop_pk(uint32_t (&pk)[16], uint32_t dbase, uint32_t w0, uint32_t p8) {
#pragma unroll
  for (int wi = 0; wi < 8; ++wi) {
    const uint32_t word = drop_mix(dbase + (w0 + wi) * DROP_C_WORD);
    pk[2 * wi] &= drop_pair_mask(word, 0, p8);
    pk[2 * wi + 1] &= drop_pair_mask(word, 1, p8);
  }
}

// DROP: the training instances that apply dropout to the probabilities -- a template parameter, not a runtime branch: with
// the branch in the loop the inference kernels ran 4 - 11 % slower (500x500 0.546 against 0.525 ms) although it was
// never taken.
template <int DH, bool PAIRED, bool DEFER, bool DROP = false>
__global__ void __launch_bounds__(A3_THREADS, 1)
attention_fwd3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                      const Attn3Params p) {
  using L = Attn3Smem<DH>;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t TILE_COLS = 256;   // per query tile: S0 at +0, S1 at +64, O at +128 (PSEP: P at +128, O at +160)
  // PSEP (long-key general form, dh <= 96): P does NOT overlay the S buffer it was computed from but lives in the 32
  // columns that are spare next to 2 x 64 (S) + dh (O).  An S buffer is then free again as soon as the warpgroup holds
  // its scores in registers (s_free), not only after PV has read P out of it: S(g+2) is issued at the START of step g
  // instead of behind PV(g), so S(j+1) has landed when step j begins and its tensor-memory load can pass under the
  // exponentials of step j (with P over S the look-ahead was less than one step: r02_attention_issue_notes.txt item 8).
  // Measured: NOT faster (500x500 0.607 ms against 0.523) -- the load is issued early, but tcgen05.wait::ld, the P-store
  // drain and the barrier waits each cost their ~200 cycles whether or not the data has long arrived, and the extra
  // s_free arrive / pv_done wait add two more such instructions per step (r02_attention_issue_notes.txt item 9).  Kept
  // for A/B runs under HRIEMO_ATTN_PSEP.
#ifdef HRIEMO_ATTN_PSEP
  constexpr bool PSEP = !PAIRED && !DEFER && DH <= 96;
#else
  constexpr bool PSEP = false;
#endif
  constexpr uint32_t P_COL = 128;
  constexpr uint32_t O_COL = PSEP ? 160 : 128;
  constexpr int KS = L::KV_STAGES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t sQ = base + L::Q_OFF, sK = base + L::K_OFF, sV = base + L::V_OFF;
  const uint32_t bars = base + L::BAR_OFF;
  const uint32_t b_qfull = bars + 0 * 8;     // [tile][buffer]
  const uint32_t b_qempty = bars + 4 * 8;    // [tile][buffer]
  const uint32_t b_kfull = bars + 8 * 8;     // [stage]
  const uint32_t b_kempty = bars + 11 * 8;
  const uint32_t b_vfull = bars + 14 * 8;
  const uint32_t b_vempty = bars + 17 * 8;
  const uint32_t b_sfull = bars + 20 * 8;    // [tile][S buffer]
  // p_full / pv_done are indexed [tile][k & 1] by the tile's running step count k: an mbarrier wait
  // only sees one parity bit, and with S issued two steps ahead the two sides may be up to two
  // steps apart; alternating barriers keep every wait at most one phase behind its barrier.
  const uint32_t b_pfull = bars + 24 * 8;
  const uint32_t b_pvdone = bars + 28 * 8;
  const uint32_t b_sfree = bars + 32 * 8;    // [tile][S buffer], PSEP form only: the warpgroup has the buffer's scores in registers
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::TMEM_SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = p.n_kv;
  // Work items (utterance, head, query-tile pair) are dealt to the CTAs either round-robin or in
  // contiguous runs (p.contiguous).  Measured on B200: short key sequences (<= 6 steps per item,
  // where the per-item work is small and consecutive items share K / V and the key mask) gain 5-9 %
  // from contiguous runs, long ones lose 3-8 % (neighbouring CTAs then no longer stream the same
  // K / V tiles at the same time), so the launcher picks by the number of key steps.
  const uint32_t n_items_all = static_cast<uint32_t>(p.n_items);
  const uint32_t per_cta = (n_items_all + gridDim.x - 1) / gridDim.x;
  const uint32_t item_first = p.contiguous ? blockIdx.x * per_cta : blockIdx.x;
  const uint32_t item_stride = p.contiguous ? 1u : gridDim.x;
  const uint32_t item_last = p.contiguous ? min(item_first + per_cta, n_items_all) : n_items_all;  // exclusive
  // key steps of a work item: all of them, or only up to the utterance's last valid key (a tile of
  // PAD keys adds exactly 0 to every row sum, so skipping it is bit-exact); every role asks the same way
  auto steps_of = [&](uint32_t item) -> int {
    if (p.kv_steps == nullptr) return n_kv;
    return __ldg(p.kv_steps + item / static_cast<uint32_t>(p.items_per_b));
  };
  constexpr bool paired = PAIRED;   // compile-time: the general form keeps its original code paths

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    for (int s = 0; s < 4; ++s) {
      mbar_init(b_qfull + s * 8, 1);
      mbar_init(b_qempty + s * 8, 4);   // released by the 4 warps of the tile's warpgroup (epilogue staging)
      mbar_init(b_sfull + s * 8, 1);
      mbar_init(b_pvdone + s * 8, 1);
      mbar_init(b_sfree + s * 8, 4);    // one arrival per softmax warp of the tile
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(b_kfull + s * 8, 1);
      mbar_init(b_kempty + s * 8, 2);   // one arrival per MMA issuer
      mbar_init(b_vfull + s * 8, 1);
      mbar_init(b_vempty + s * 8, 2);
    }
#ifdef HRIEMO_ATTN_V3_SOFTMAX
    for (int s = 0; s < 4; ++s) mbar_init(b_pfull + s * 8, 128);
#else
    for (int s = 0; s < 4; ++s) mbar_init(b_pfull + s * 8, 4);   // one arrival per softmax warp (lane 0, after __syncwarp)
#endif
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(base + L::TMEM_SLOT_OFF);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // elect.sync, not lane == 0: cp.async.bulk.tensor takes uniform-register operands (see the issuers)
      uint32_t qcnt[2] = {0, 0};  // Q loads issued per query tile -> buffer and phase
      if constexpr (!PAIRED) {
        // K(g) then V(g), step by step; Q tiles at the start of an item
        uint32_t g = 0;             // flat step index -> K/V ring stage and phase
        for (uint32_t item = item_first; item < item_last; item += item_stride) {
          const uint32_t bh = item / static_cast<uint32_t>(p.n_qp);
          const int qp = static_cast<int>(item - bh * static_cast<uint32_t>(p.n_qp));
          const int b = static_cast<int>(bh / static_cast<uint32_t>(p.H));
          const int h = static_cast<int>(bh - static_cast<uint32_t>(b) * static_cast<uint32_t>(p.H));
          const int q0 = qp * 2 * A3_BQ;
          for (int t = 0; t < 2; ++t) {
            if (q0 + t * A3_BQ >= p.Tq) break;
            const uint32_t qb = qcnt[t] & 1u, qpar = (qcnt[t] >> 1) & 1u;
            const uint32_t bar = (t * 2 + qb) * 8;
            mbar_wait(b_qempty + bar, qpar ^ 1);
            mbar_arrive_expect_tx(b_qfull + bar, L::Q_TILE);
            for (int c = 0; c < L::QCH; ++c)
              tma_load_2d(&tm_q, b_qfull + bar, sQ + (t * 2 + qb) * L::Q_TILE + c * L::Q_CHUNK, h * DH + c * 64,
                          b * p.Tq + q0 + t * A3_BQ);
            ++qcnt[t];
          }
          const int nk = steps_of(item);
          for (int j = 0; j < nk; ++j, ++g) {
            const uint32_t s = g % KS, par = (g / KS) & 1u;
            mbar_wait(b_kempty + s * 8, par ^ 1);
            mbar_arrive_expect_tx(b_kfull + s * 8, L::K_STAGE);
            for (int c = 0; c < L::QCH; ++c)
              tma_load_2d(&tm_k, b_kfull + s * 8, sK + s * L::K_STAGE + c * L::K_CHUNK, h * DH + c * 64,
                          b * p.Tk + j * A3_BKV);
            mbar_wait(b_vempty + s * 8, par ^ 1);
            mbar_arrive_expect_tx(b_vfull + s * 8, L::V_STAGE);
            for (int c = 0; c < DH / 32; ++c)
              tma_load_3d(&tm_v, b_vfull + s * 8, sV + s * L::V_STAGE + c * L::V_GROUP, h * DH + c * 32,
                          j * A3_BKV, b);
          }
        }
      } else {
        // Paired-head mode: two cursors over the flat step sequence, the K cursor (which also brings in the Q
        // tiles at the start of an item) and the V cursor.  The K cursor runs up to two flat steps ahead of the
        // V cursor WITHOUT ever blocking while it is ahead -- a V stage is held until its PV has retired (a
        // whole softmax step), K stages are free again as soon as S has been issued, and the S look-ahead of
        // both tiles needs K(g+2), K(g+3) long before V(g+1)'s slot frees (with K and V in lock-step the
        // 64 x 500 shape gained nothing from the second tile).
        struct Cur { uint32_t item, g; int f, nk, b, h; };
        auto cur_item = [&](Cur& c) {   // (utterance, first head of the pair) of the cursor's item
          c.nk = c.item < item_last ? steps_of(c.item) : 0;
          c.b = static_cast<int>(c.item / static_cast<uint32_t>(p.items_per_b));
          c.h = static_cast<int>(c.item - static_cast<uint32_t>(c.b) * static_cast<uint32_t>(p.items_per_b)) * 2;
        };
        auto cur_next = [&](Cur& c) {
          ++c.g;
          if (++c.f == 2 * c.nk) {
            c.f = 0;
            c.item += item_stride;
            cur_item(c);
          }
        };
        // Q tiles (first flat step of an item) + K tile of the cursor's step.  Non-blocking form: returns
        // false, with nothing issued, unless every buffer the step needs is free.
        auto k_issue = [&](Cur& c, bool blocking) -> bool {
          const uint32_t s = c.g % KS, par = (c.g / KS) & 1u;
          if (c.f == 0) {
            for (int t = 0; t < 2; ++t) {
              const uint32_t bar = (t * 2 + (qcnt[t] & 1u)) * 8, qpar = (qcnt[t] >> 1) & 1u;
              if (blocking) mbar_wait(b_qempty + bar, qpar ^ 1);
              else if (!mbar_test_wait(b_qempty + bar, qpar ^ 1)) return false;
            }
          }
          if (blocking) mbar_wait(b_kempty + s * 8, par ^ 1);
          else if (!mbar_test_wait(b_kempty + s * 8, par ^ 1)) return false;
          if (c.f == 0) {
            for (int t = 0; t < 2; ++t) {
              const uint32_t qb = qcnt[t] & 1u;
              const uint32_t bar = (t * 2 + qb) * 8;
              mbar_arrive_expect_tx(b_qfull + bar, L::Q_TILE);
              for (int cc = 0; cc < L::QCH; ++cc)
                tma_load_2d(&tm_q, b_qfull + bar, sQ + (t * 2 + qb) * L::Q_TILE + cc * L::Q_CHUNK,
                            (c.h + t) * DH + cc * 64, c.b * p.Tq);
              ++qcnt[t];
            }
          }
          const int hk = c.h + (c.f & 1);   // head whose K / V this flat step carries; key step f >> 1
          mbar_arrive_expect_tx(b_kfull + s * 8, L::K_STAGE);
          for (int cc = 0; cc < L::QCH; ++cc)
            tma_load_2d(&tm_k, b_kfull + s * 8, sK + s * L::K_STAGE + cc * L::K_CHUNK, hk * DH + cc * 64,
                        c.b * p.Tk + (c.f >> 1) * A3_BKV);
          cur_next(c);
          return true;
        };
        auto v_issue = [&](Cur& c) {
          const uint32_t s = c.g % KS, par = (c.g / KS) & 1u;
          const int hk = c.h + (c.f & 1);
          mbar_wait(b_vempty + s * 8, par ^ 1);
          mbar_arrive_expect_tx(b_vfull + s * 8, L::V_STAGE);
          for (int cc = 0; cc < DH / 32; ++cc)
            tma_load_3d(&tm_v, b_vfull + s * 8, sV + s * L::V_STAGE + cc * L::V_GROUP, hk * DH + cc * 32,
                        (c.f >> 1) * A3_BKV, c.b);
          cur_next(c);
        };
        Cur kc;
        kc.item = item_first; kc.g = 0; kc.f = 0;
        cur_item(kc);
        Cur vc = kc;
        while (vc.item < item_last) {
          while (kc.item < item_last && kc.g <= vc.g + 2u) {
            if (!k_issue(kc, /*blocking=*/kc.g == vc.g)) break;   // K(g) must be out before V(g); beyond that, opportunistic
          }
          v_issue(vc);
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===================== MMA issuers: warp 1 drives query tile 0, warp 10 query tile 1 =====
    // Each walks the CTA's flat step sequence for its own tile: PV_t(g) as soon as P_t(g) is
    // there, then S_t(g+2) into the S buffer PV_t(g) just consumed (same thread => in order).
    // The two tiles only meet at the shared K / V^T stages, whose "empty" barriers expect one
    // arrival from each issuer (a plain arrive when the tile does not exist for an item).
    //
    // ONE lane runs the loop, and it is chosen with elect.sync, not `lane == 0`: tcgen05.mma / commit take their
    // descriptors in UNIFORM registers, and only behind elect.sync does the compiler know that a single lane is
    // active.  Behind `if (lane == 0)` it wrapped every MMA in an elect / 3 x R2UR.BROADCAST / branch-back loop
    // (for lanes that might hold different values) -- ~75 cycles per instruction by the pipeline trace, 1 100 cycles
    // of issuing per 64-key step for 384 cycles of tensor work, which made the ISSUER the critical path of the
    // kernel (the softmax warpgroups sat waiting for S: the v4 softmax loop alone changed nothing).  The tile index
    // is a compile-time constant (one copy of the loop per issuer warp) and the tensor-memory base is the
    // constant 0, so the descriptors are formed by the uniform datapath and the MMAs issue back to back.
    auto run_issuer = [&](auto tile_c) {
      constexpr int t = decltype(tile_c)::value;
      // This CTA allocates all 512 tensor-memory columns of its SM, so the allocation starts at column 0, lane 0:
      // the base is the CONSTANT 0 (checked), not a value loaded from shared memory.
      if (tmem_base != 0u) __trap();
      constexpr uint32_t tmem_base_u = 0u;
      auto wait_u = [&](uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); };
      auto test_u = [&](uint32_t bar, uint32_t parity) -> bool { return mbar_test_wait(bar, parity); };
      auto steps_u = [&](uint32_t item) -> int { return steps_of(item); };
      constexpr bool leader = true;
      constexpr uint32_t idesc_s = umma_idesc_bf16(A3_BQ, A3_BKV);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(A3_BQ, DH) | kUmmaBMajorMN;
      const uint32_t tile_tmem = tmem_base_u + t * TILE_COLS;
      const uint64_t k_desc0 = umma_desc_sw128(sK);
      const uint64_t v_desc0 = umma_desc_mn_sw64(sV, L::V_GROUP);
      // does this tile exist for the item (general mode; in paired-head mode both tiles always do)
      auto tile_active = [&](uint32_t item) {
        return paired || t == 0 ||
               static_cast<int>(item % static_cast<uint32_t>(p.n_qp)) * 2 * A3_BQ + A3_BQ < p.Tq;
      };
      // flat steps of an item, and which of them belong to this tile (paired-head mode: every second one)
      auto flat_of = [&](int nk) { return paired ? 2 * nk : nk; };
      auto mine = [&](int f) { return !paired || (f & 1) == t; };
      auto key_step = [&](int f) { return paired ? f >> 1 : f; };
      // S buffer of a flat step: by the tile's own step count (flat steps per item are even in paired-head mode)
      auto sbuf_of = [&](uint32_t g) { return paired ? (g >> 1) & 1u : g & 1u; };
      // ---- S cursor (two flat steps ahead of the PV cursor)
      uint32_t s_g = 0, s_item = item_first, qcnt = 0;
      [[maybe_unused]] uint32_t s_uses[2] = {0u, 0u};   // PSEP: S tiles of this query tile issued into each buffer so far
      int s_f = 0;
      int s_nk = s_item < item_last ? steps_u(s_item) : 0;
      bool s_act = tile_active(s_item);
      // S of the cursor's flat step.  Non-blocking form: returns false, with nothing issued, unless the Q tile
      // (first key step of an item) and the K tile have landed.
      auto issue_s = [&](bool blocking) -> bool {
        const uint32_t ks = s_g % KS, kpar = (s_g / KS) & 1u;
        const uint32_t qslot = t * 2 + (qcnt & 1u);
        const bool act = s_act && mine(s_f);
        const int s_j = key_step(s_f);
        if (leader) ATRACE(t, s_g, 4);
        if (blocking) {
          if (act && s_j == 0) wait_u(b_qfull + qslot * 8, (qcnt >> 1) & 1u);
          wait_u(b_kfull + ks * 8, kpar);
        } else {
          // only the Q tile can be far away (it is loaded once the warpgroup has freed its buffer); a K tile is
          // in flight, and waiting for it here is what keeps S(g+2) right behind PV(g) on the long-key shapes
          if (act && s_j == 0 && !test_u(b_qfull + qslot * 8, (qcnt >> 1) & 1u)) return false;
          wait_u(b_kfull + ks * 8, kpar);
        }
        if (leader) ATRACE(t, s_g, 5);
        if (act) {
          if constexpr (PSEP) {   // the buffer's previous scores must be in the warpgroup's registers
            const uint32_t sb = sbuf_of(s_g);
            if (s_uses[sb] > 0) wait_u(b_sfree + (t * 2 + sb) * 8, (s_uses[sb] - 1u) & 1u);
            ++s_uses[sb];
          }
          tc_fence_after_sync();
          const uint64_t q_desc = umma_desc_sw128(sQ + qslot * L::Q_TILE);
          const uint64_t k_desc = k_desc0 + ((ks * L::K_STAGE) >> 4);
          const uint32_t d_tmem = tile_tmem + sbuf_of(s_g) * A3_BKV;
          {
#pragma unroll
            for (int st = 0; st < DH / 16; ++st) {
              umma_bf16(d_tmem, q_desc + (((st >> 2) * L::Q_CHUNK + (st & 3) * 32) >> 4),
                        k_desc + (((st >> 2) * L::K_CHUNK + (st & 3) * 32) >> 4), idesc_s, st != 0);
            }
            umma_commit(b_sfull + (t * 2 + sbuf_of(s_g)) * 8);
            umma_commit(b_kempty + ks * 8);
            ATRACE(t, s_g, 6);
          }
          if (s_j == s_nk - 1) ++qcnt;  // the Q buffer is released by the warpgroup after its epilogue
        } else if (leader) {
          mbar_arrive(b_kempty + ks * 8);
        }
        ++s_g;
        if (++s_f == flat_of(s_nk)) {
          s_f = 0;
          s_item += item_stride;
          s_act = s_item < item_last && tile_active(s_item);
          s_nk = s_item < item_last ? steps_u(s_item) : 0;
        }
        return true;
      };
      // General mode: the S cursor leads the PV cursor by two steps (S(g+2) goes into the buffer whose P the
      // PV(g) just issued consumed).  Paired-head mode: by THREE flat steps, so that the tile's next own step
      // S(g+2) is issued one iteration early, at the other tile's flat step g-1 (where this issuer only passes
      // the stage on): it lands in the buffer of P(g-2), whose PV was issued at iteration g-2 -- the same
      // one-step look-ahead for the softmax as in general mode.
      //
      // DEFER form (items of one or two key steps, chosen by the launcher): the look-ahead never holds up a PV.
      // S steps up to `s_allowed` are issued as soon as their Q tile has landed, polled while the issuer waits
      // for V or P.  Issued unconditionally after each PV -- the other form, right for long items -- S(g+2) of a
      // one-key-step item sat 4 500 - 7 600 cycles on the Q tile of the item two ahead (which is only loaded once
      // the warpgroup frees the buffer at the start of item g+1), and the PV of item g+1, whose P had been ready
      // for 3 000 cycles, waited behind it: profiles/r01_attention_v12_pipeline_trace.txt.  The polling form is
      // not used for long items: an issuer that polls shares its scheduler with two softmax warps (-4..8 %).
      constexpr uint32_t S_LEAD = paired ? 3u : 2u;
      uint32_t s_allowed = S_LEAD;
      auto pump = [&](bool blocking) {
        while (s_item < item_last && s_g < s_allowed)
          if (!issue_s(blocking)) break;
      };
      if constexpr (DEFER) {
        pump(true);
      } else {
        if (s_item < item_last) issue_s(true);
        if (s_item < item_last) issue_s(true);
        if (paired && s_item < item_last) issue_s(true);
      }
      auto wait_pumping = [&](uint32_t bar, uint32_t parity) {
        if constexpr (DEFER) {
          while (s_item < item_last && s_g < s_allowed) {
            if (test_u(bar, parity)) return;
            issue_s(false);
          }
        }
        wait_u(bar, parity);
      };
      // ---- PV cursor
      uint32_t pcnt = 0, pv_item = item_first;
      int pv_f = 0;
      int pv_nk = pv_item < item_last ? steps_u(pv_item) : 0;
      bool pv_act = tile_active(pv_item);
      for (uint32_t g = 0; pv_item < item_last; ++g) {
        const uint32_t vs = g % KS, vpar = (g / KS) & 1u;
        if constexpr (PSEP) {
          if (s_item < item_last) issue_s(true);   // S(g+2) as soon as the warpgroup has loaded S(g): before PV(g), not behind it
        }
        if (leader) ATRACE(t, g, 0);
        wait_pumping(b_vfull + vs * 8, vpar);   // never wait without moving the S cursor
        if (leader) ATRACE(t, g, 1);
        if (pv_act && mine(pv_f)) {
          const int pv_j = key_step(pv_f);
          const int rem = p.Tk - pv_j * A3_BKV;  // keys left from this step on (> 0)
          const uint32_t slot = t * 2 + (pcnt & 1u);
          wait_pumping(b_pfull + slot * 8, (pcnt >> 1) & 1u);   // P(g) needs S(g): keep the S cursor moving
          if (leader) ATRACE(t, g, 2);
          ++pcnt;
          tc_fence_after_sync();
          const uint64_t v_desc = v_desc0 + ((vs * L::V_STAGE) >> 4);
          const uint32_t p_tmem = PSEP ? tile_tmem + P_COL : tile_tmem + sbuf_of(g) * A3_BKV;
          {
#pragma unroll
            for (int st = 0; st < A3_BKV / 16; ++st) {
              if (st * 16 < rem)  // P is zero beyond Tk: skip those K-steps (16 keys = 16 rows of 64 B)
                umma_bf16_ts(tile_tmem + O_COL, p_tmem + st * 8, v_desc + ((st * 16 * 64) >> 4), idesc_pv,
                             (pv_j | st) != 0);
            }
            umma_commit(b_pvdone + slot * 8);
            umma_commit(b_vempty + vs * 8);
            ATRACE(t, g, 3);
          }
        } else if (leader) {
          mbar_arrive(b_vempty + vs * 8);
        }
        if constexpr (DEFER) {
          s_allowed = g + 1 + S_LEAD;
          pump(false);
        } else if constexpr (!PSEP) {
          if (s_item < item_last) issue_s(true);  // general: S(g+2) reuses the S buffer whose P was just consumed
        }
        if (++pv_f == flat_of(pv_nk)) {
          pv_f = 0;
          pv_item += item_stride;
          pv_act = pv_item < item_last && tile_active(pv_item);
          pv_nk = pv_item < item_last ? steps_u(pv_item) : 0;
        }
      }
    };
    if (elect_one()) {
      if (warp == 1) run_issuer(std::integral_constant<int, 0>{});
      else run_issuer(std::integral_constant<int, 1>{});
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int wg = (warp - 2) >> 2;          // query tile handled by this warpgroup
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
      const int wg_tid = threadIdx.x - 64 - wg * 128;
    float* caps = reinterpret_cast<float*>(base_ptr + L::DYN_OFF) + wg * n_kv * A3_BKV;
    int* flags = reinterpret_cast<int*>(base_ptr + L::DYN_OFF + 2 * n_kv * A3_BKV * 4) + wg * n_kv;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_tile = tmem_base + lane_sel + wg * TILE_COLS;
    const uint32_t t_o = t_tile + O_COL;
    const float sc = p.scale_log2;
    uint32_t scnt0 = 0, scnt1 = 0;   // completed uses of each S buffer of this tile
    // PV(k) of this tile commits to pv_done[k & 1].  Waiting for PV(k) is only safe while PV(k+2)
    // cannot have completed (mbarrier waits see one parity bit), i.e. before P(k+2) is handed over:
    // the consumer may lag the producer by at most two tiles, and never does more.
    uint32_t pv_seen = 0;        // PVs of this tile known to have retired
    uint32_t pv_issued = 0;      // P tiles handed to the MMA warp so far
    auto consume_pv = [&](uint32_t target) {
      while (pv_seen < target) {
        mbar_wait(b_pvdone + (wg * 2 + (pv_seen & 1u)) * 8, (pv_seen >> 1) & 1u);
        ++pv_seen;
      }
    };
    uint32_t g = 0;              // flat step index of this CTA at the start of the current item
    int pending_qslot = -1;      // Q buffer whose O store has been issued but not yet waited for
    auto release_q = [&]() {
      if (pending_qslot >= 0) {
        if (lane == 0) {
          bulk_wait_read<0>();   // staging bytes have been read: the Q buffer may be refilled
          mbar_arrive(b_qempty + pending_qslot * 8);
        }
        pending_qslot = -1;
      }
    };
#ifndef HRIEMO_ATTN_NO_PINGPONG
    if (wg == 1) asm volatile("bar.arrive %0, 256;" ::"r"(3) : "memory");   // warpgroup 0 takes the first turn
#endif
    uint32_t qcnt_w = 0;         // items this tile has processed -> which Q buffer it used
    int cur_b = -1;

    int nk = 0;
    for (uint32_t item = item_first; item < item_last; item += item_stride, g += paired ? 2 * nk : nk) {
      nk = steps_of(item);
      const int b = static_cast<int>(item / static_cast<uint32_t>(p.items_per_b));
      const uint32_t in_b = item - static_cast<uint32_t>(b) * static_cast<uint32_t>(p.items_per_b);
      const int h = paired ? static_cast<int>(in_b) * 2 + wg : static_cast<int>(in_b / static_cast<uint32_t>(p.n_qp));
      const int q0 = paired ? 0 : static_cast<int>(in_b % static_cast<uint32_t>(p.n_qp)) * 2 * A3_BQ + wg * A3_BQ;
      // dropout on the probabilities (training): hash base of this thread's query row in the (utterance, head) stream
      [[maybe_unused]] const uint32_t dbase =
          drop_key_bh(p.drop_key, static_cast<uint32_t>(b * p.H + h)) + static_cast<uint32_t>(q0 + quad * 32 + lane) * DROP_C_ROW;
      if (q0 >= p.Tq) {  // this warpgroup's tile does not exist for this item: only keep the exp turn-taking alive
#ifndef HRIEMO_ATTN_NO_PINGPONG
        for (int j = 0; j < nk; ++j) {
          asm volatile("bar.sync %0, 256;" ::"r"(3 + wg) : "memory");
          asm volatile("bar.arrive %0, 256;" ::"r"(4 - wg) : "memory");
        }
#endif
        continue;
      }

      if (cur_b < 0 || (b != cur_b && p.key_pad != nullptr)) {   // without a mask the caps only depend on Tk
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        for (int j = wg_tid; j < n_kv; j += 128) flags[j] = 0;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        for (int kk = wg_tid; kk < n_kv * A3_BKV; kk += 128) {
          bool pad = kk >= p.Tk;
          if (!pad && p.key_pad != nullptr) pad = p.key_pad[static_cast<int64_t>(b) * p.Tk + kk] != 0;
          caps[kk] = pad ? -INFINITY : INFINITY;
          if (pad) flags[kk / A3_BKV] = 1;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        cur_b = b;
      }

#ifdef HRIEMO_ATTN_V3_SOFTMAX
      if (nk <= 2) release_q();   // short items: the producer needs the buffer back sooner (two items ahead)
      float m_run = -INFINITY;  // running reference maximum, in log2 units (score * scale * log2 e)
      float l_run = 0.0f;
      for (int j = 0; j < nk; ++j) {
        // flat step of this tile's j-th key step: g + j, or g + 2 j + wg in paired-head mode (g is even there)
        const uint32_t sbuf = paired ? ((g >> 1) + static_cast<uint32_t>(j)) & 1u : (g + static_cast<uint32_t>(j)) & 1u;
        const uint32_t t_s = t_tile + sbuf * A3_BKV;
        const int rem = p.Tk - j * A3_BKV;
        const bool two = rem > 32;                   // second 32-key chunk holds a valid key
        const bool masked = flags[j] != 0;           // warp-uniform
        const float* cap_j = caps + j * A3_BKV;
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 0);
        mbar_wait(b_sfull + (wg * 2 + sbuf) * 8, (sbuf ? scnt1 : scnt0) & 1u);
        if (sbuf) ++scnt1; else ++scnt0;
        tc_fence_after_sync();
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 1);

        uint32_t va[32], vb[32];
        tmem_ld32(t_s, va);
        if (two) tmem_ld32(t_s + 32, vb);
        tmem_ld_wait();
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 2);
        if (masked) {
          apply_caps(va, cap_j);
          if (two) apply_caps(vb, cap_j + 32);
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        max4(m4, va);
        if (two) max4(m4, vb);
        const float tile_max = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;  // sc > 0

        // ---- lazy running maximum; O / l rescale only when the maximum grew by more than 2^TAU
        if (j == 0) {
          m_run = tile_max;
        } else {
          const bool need = tile_max > m_run + A3_LAZY_TAU;
          const bool any_need = __any_sync(0xffffffffu, need);
          // all but the latest PV retired long ago; the latest one is awaited only for a rescale
          consume_pv(any_need ? pv_issued : pv_issued - 1);
          if (any_need) {
            tc_fence_after_sync();
            const float alpha = need ? ex2_approx(m_run - tile_max) : 1.0f;
            if (need) m_run = tile_max;
            l_run *= alpha;
#pragma unroll 1
            for (int c = 0; c < DH / 32; ++c) {
              uint32_t v[32];
              tmem_ld32(t_o + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st32(t_o + c * 32, v);
            }
          }
        }
        const float neg_m = (m_run == -INFINITY) ? 0.0f : -m_run;
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 3);

        // ---- p = 2^(s*scale - m), row sum, bf16 P into the first 32 columns of this S buffer.
        // The two warpgroups take turns here (named barriers 3 / 4): left alone they run in lock-step
        // and their exponentials collide on the quarter-rate XU pipe (profiles/r01_attention_v8_pipeline_trace);
        // in turns, one warpgroup's exp phase overlaps the other's waits, loads and row maxima.
#ifndef HRIEMO_ATTN_NO_PINGPONG
        asm volatile("bar.sync %0, 256;" ::"r"(3 + wg) : "memory");
#endif
        float l4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        uint32_t pk[16];
        exp_pack(va, sc, neg_m, l4, pk);
        tmem_st16(t_s, pk);
        if (two) {
          exp_pack(vb, sc, neg_m, l4, pk);
          tmem_st16(t_s + 16, pk);
        }
        l_run += (l4[0] + l4[1]) + (l4[2] + l4[3]);
#ifndef HRIEMO_ATTN_NO_PINGPONG
        asm volatile("bar.arrive %0, 256;" ::"r"(4 - wg) : "memory");   // the other warpgroup's turn
#endif
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 4);
        tmem_st_wait();
        tc_fence_before_sync();
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 5);
        mbar_arrive(b_pfull + (wg * 2 + (pv_issued & 1u)) * 8);
        ++pv_issued;
        if (j == 0 && nk > 2) release_q();
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 6);
      }

#elif !defined(HRIEMO_ATTN_V5_SOFTMAX)
      // ---- v4 softmax loop (the default).  Against the first form (kept under HRIEMO_ATTN_V3_SOFTMAX for A/B runs):
      //  * the scores of step j+1 are fetched (wait S, tcgen05.ld issued) while the P stores of step j drain, so the
      //    fixed latencies of a step (barrier wait, tensor-memory load, store drain, hand-off) overlap instead of adding up;
      //  * no barrier wait per step for the PV two steps back: S(k) is issued behind PV(k-2) by the same thread, so
      //    S(k) landed implies PV(k-2) retired, and PV(k-1) is only awaited when a rescale or the epilogue needs O;
      //  * P is handed over with one mbarrier arrival per warp.
      if (nk <= 2) release_q();   // short items: the producer needs the buffer back sooner (two items ahead)
      float m_run = -INFINITY;  // running reference maximum, in log2 units (score * scale * log2 e)
      float l_run = 0.0f;
      uint32_t va[32], vb[32];
      if constexpr (PSEP) {
        // ---- PSEP loop (see the constant's comment): P in its own 32 columns, S buffers released as soon as their scores
        // are in registers, exact row maximum first, the next step's scores loaded under this step's exponentials.
        const uint32_t t_p = t_tile + P_COL;
        {  // the item's first scores
          const uint32_t sbuf0 = g & 1u;
          if (wg_tid == 0) ATRACE(2 + wg, g, 0);
          mbar_wait(b_sfull + (wg * 2 + sbuf0) * 8, (sbuf0 ? scnt1 : scnt0) & 1u);
          if (sbuf0) ++scnt1; else ++scnt0;
          tc_fence_after_sync();
          if (wg_tid == 0) ATRACE(2 + wg, g, 1);
          tmem_ld32(t_tile + sbuf0 * A3_BKV, va);
          if (p.Tk > 32) tmem_ld32(t_tile + sbuf0 * A3_BKV + 32, vb);
        }
        for (int j = 0; j < nk; ++j) {
          const uint32_t sbuf = (g + static_cast<uint32_t>(j)) & 1u;
          const int rem = p.Tk - j * A3_BKV;
          const bool two = rem > 32;                   // second 32-key chunk holds a valid key
          const bool masked = flags[j] != 0;           // warp-uniform
          const float* cap_j = caps + j * A3_BKV;
          tmem_ld_wait();
          // the scores are in registers: the buffer may take S(j+2)
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(b_sfree + (wg * 2 + sbuf) * 8);
          if (wg_tid == 0) ATRACE(2 + wg, g + j, 2);
          if (masked) {
            apply_caps(va, cap_j);
            if (two) apply_caps(vb, cap_j + 32);
          }
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          max4(m4, va);
          if (two) max4(m4, vb);
          const float tile_max = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;  // sc > 0
          // P has ONE buffer: PV(j-1) must have read it before this step's probabilities go in (it was issued a whole
          // step ago; every PV is awaited in order, so no wait can be answered by a stale phase)
          consume_pv(pv_issued);
          if (j == 0) {
            m_run = tile_max;
          } else {
            const bool need = tile_max > m_run + A3_LAZY_TAU;
            if (__any_sync(0xffffffffu, need)) {   // rare: O / l rescale (every PV of this tile has retired, see above)
              tc_fence_after_sync();
              const float alpha = need ? ex2_approx(m_run - tile_max) : 1.0f;
              if (need) m_run = tile_max;
              l_run *= alpha;
#pragma unroll 1
              for (int c = 0; c < DH / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(t_o + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                tmem_st32(t_o + c * 32, v);
              }
            }
          }
          const float neg_m = (m_run == -INFINITY) ? 0.0f : -m_run;
          if (wg_tid == 0) ATRACE(2 + wg, g + j, 3);
#ifndef HRIEMO_ATTN_NO_PINGPONG
          asm volatile("bar.sync %0, 256;" ::"r"(3 + wg) : "memory");
#endif
          float l4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          uint32_t pk[16];
          float2 xs[16];
          scale32(va, sc, neg_m, xs);
          exp_pack_x(xs, l4, pk);
          tmem_st16(t_p, pk);
          if (two) scale32(vb, sc, neg_m, xs);     // va and vb are dead from here on
          if (j + 1 < nk) {   // S(j+1) was issued at the start of step j-1: it has landed
            const uint32_t nbuf = sbuf ^ 1u;
            mbar_wait(b_sfull + (wg * 2 + nbuf) * 8, (nbuf ? scnt1 : scnt0) & 1u);
            if (nbuf) ++scnt1; else ++scnt0;
            tc_fence_after_sync();
            tmem_ld32(t_tile + nbuf * A3_BKV, va);
            if (p.Tk - (j + 1) * A3_BKV > 32) tmem_ld32(t_tile + nbuf * A3_BKV + 32, vb);
          }
          if (two) {
            exp_pack_x(xs, l4, pk);
            tmem_st16(t_p + 16, pk);
          }
#ifndef HRIEMO_ATTN_NO_PINGPONG
          asm volatile("bar.arrive %0, 256;" ::"r"(4 - wg) : "memory");   // the other warpgroup's turn
#endif
          l_run += (l4[0] + l4[1]) + (l4[2] + l4[3]);
          if (wg_tid == 0) ATRACE(2 + wg, g + j, 4);
          tmem_st_wait();
          tc_fence_before_sync();
          if (wg_tid == 0) ATRACE(2 + wg, g + j, 5);
          __syncwarp();
          if (lane == 0) mbar_arrive(b_pfull + (wg * 2 + (pv_issued & 1u)) * 8);
          ++pv_issued;
          if (j == 0 && nk > 2) release_q();
          if (wg_tid == 0) ATRACE(2 + wg, g + j, 6);
        }
      } else {
      // the item's first scores
      {
        const uint32_t sbuf0 = paired ? (g >> 1) & 1u : g & 1u;
        if (wg_tid == 0) ATRACE(2 + wg, g, 0);
        mbar_wait(b_sfull + (wg * 2 + sbuf0) * 8, (sbuf0 ? scnt1 : scnt0) & 1u);
        if (sbuf0) ++scnt1; else ++scnt0;
        tc_fence_after_sync();
        if (wg_tid == 0) ATRACE(2 + wg, g, 1);
        tmem_ld32(t_tile + sbuf0 * A3_BKV, va);
        if (p.Tk > 32) tmem_ld32(t_tile + sbuf0 * A3_BKV + 32, vb);
      }
      for (int j = 0; j < nk; ++j) {
        // flat step of this tile's j-th key step: g + j, or g + 2 j + wg in paired-head mode (g is even there)
        const uint32_t sbuf = paired ? ((g >> 1) + static_cast<uint32_t>(j)) & 1u : (g + static_cast<uint32_t>(j)) & 1u;
        const uint32_t t_s = t_tile + sbuf * A3_BKV;
        const int rem = p.Tk - j * A3_BKV;
        const bool two = rem > 32;                   // second 32-key chunk holds a valid key
        const bool masked = flags[j] != 0;           // warp-uniform
        const float* cap_j = caps + j * A3_BKV;
        tmem_ld_wait();
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 2);
        if (masked) {
          apply_caps(va, cap_j);
          if (two) apply_caps(vb, cap_j + 32);
        }
        // ---- row maximum of this step.  Only the item's FIRST step needs it before the exponentials (it sets the
        // reference maximum).  Later steps keep the reference unless the maximum grew by more than 2^TAU (lazy
        // rescale), which is rare: they compute p = 2^(s - m_run) with the reference they have (SPECULATIVELY) while
        // the maximum is formed on the ALU pipe next to the exponentials on the XU pipe, and check afterwards; the
        // few steps that do need a rescale repair O / l and redo their exponentials before P is handed over.
        auto row_max = [&]() {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          max4(m4, va);
          if (two) max4(m4, vb);
          return fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;  // sc > 0
        };
        if (j == 0) m_run = row_max();     // (a real branch, j is warp-uniform: later steps must not wait for their maximum)
        float neg_m = (m_run == -INFINITY) ? 0.0f : -m_run;
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 3);

        // ---- p = 2^(s*scale - m), row sum, bf16 P into the first 32 columns of this S buffer; the two warpgroups
        // take turns (named barriers 3 / 4) so that their exponentials do not collide on the quarter-rate XU pipe
#ifndef HRIEMO_ATTN_NO_PINGPONG
        asm volatile("bar.sync %0, 256;" ::"r"(3 + wg) : "memory");
#endif
        float l4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        uint32_t pk[16];
        exp_pack(va, sc, neg_m, l4, pk);
        if constexpr (DROP) drop_pk(pk, dbase, static_cast<uint32_t>(j) * 16u, p.drop_p8);
        tmem_st16(t_s, pk);
        if (two) {
          exp_pack(vb, sc, neg_m, l4, pk);
          if constexpr (DROP) drop_pk(pk, dbase, static_cast<uint32_t>(j) * 16u + 8u, p.drop_p8);
          tmem_st16(t_s + 16, pk);
        }
#ifndef HRIEMO_ATTN_NO_PINGPONG
        asm volatile("bar.arrive %0, 256;" ::"r"(4 - wg) : "memory");   // the other warpgroup's turn
#endif
        if (j > 0) {
          const float tile_max = row_max();
          const bool need = tile_max > m_run + A3_LAZY_TAU;
          if (__any_sync(0xffffffffu, need)) {
            // every PV of this tile must have retired: the latest one is PV(pv_issued - 1); the one before retired
            // before this step's S landed (see above), so this wait can not be answered by a stale phase
            const uint32_t kk = pv_issued - 1u;
            mbar_wait(b_pvdone + (wg * 2 + (kk & 1u)) * 8, (kk >> 1) & 1u);
            tc_fence_after_sync();
            const float alpha = need ? ex2_approx(m_run - tile_max) : 1.0f;
            if (need) m_run = tile_max;
            l_run *= alpha;
#pragma unroll 1
            for (int c = 0; c < DH / 32; ++c) {
              uint32_t v[32];
              tmem_ld32(t_o + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st32(t_o + c * 32, v);
            }
            // this step's exponentials again, against the new reference (the scores are still in registers)
            neg_m = (m_run == -INFINITY) ? 0.0f : -m_run;
            l4[0] = l4[1] = l4[2] = l4[3] = 0.0f;
            exp_pack(va, sc, neg_m, l4, pk);
            if constexpr (DROP) drop_pk(pk, dbase, static_cast<uint32_t>(j) * 16u, p.drop_p8);
            tmem_st16(t_s, pk);
            if (two) {
              exp_pack(vb, sc, neg_m, l4, pk);
              if constexpr (DROP) drop_pk(pk, dbase, static_cast<uint32_t>(j) * 16u + 8u, p.drop_p8);
              tmem_st16(t_s + 16, pk);
            }
          }
        }
        l_run += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 4);
        // ---- the next step's scores: wait for S(j+1) and issue its loads while the P stores drain
        if (j + 1 < nk) {
          const uint32_t nbuf = sbuf ^ 1u;
          mbar_wait(b_sfull + (wg * 2 + nbuf) * 8, (nbuf ? scnt1 : scnt0) & 1u);
          if (nbuf) ++scnt1; else ++scnt0;
          tc_fence_after_sync();
          tmem_ld32(t_tile + nbuf * A3_BKV, va);
          if (p.Tk - (j + 1) * A3_BKV > 32) tmem_ld32(t_tile + nbuf * A3_BKV + 32, vb);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 5);
        __syncwarp();
        if (lane == 0) mbar_arrive(b_pfull + (wg * 2 + (pv_issued & 1u)) * 8);
        ++pv_issued;
        if (j == 0 && nk > 2) release_q();
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 6);
      }
      }   // !PSEP
#else
      // ---- v5 softmax loop (HRIEMO_ATTN_V5_SOFTMAX; an experiment that did NOT pay: 500x500 0.553 ms against v4's 0.525 on
      // the same box, profiles/r02_attention_issue_notes.txt item 7 -- S(j+1) is issued ~600 cycles behind the hand-off of
      // P(j-1) and shares the tensor pipe with the other tile, so it has rarely landed in the middle of step j, and the
      // exact row maximum in front of the exponentials costs what the earlier load saves).  ncu's warp-state samples of v4 (profiles/r02_ncu_attention_v16_summary.txt) put only 34 %
      // of a softmax warp's time in the exponentials; 15 % went to waiting for the step's tensor-memory load at the top
      // of the loop and 8 % to the row maximum formed after the exponentials.  Here
      //  * the row maximum comes FIRST (ALU pipe, ~80 cycles) -- exact, so there is no speculative pass to redo and the
      //    raw scores are dead as soon as they have been scaled;
      //  * the second 32-score chunk is scaled before its exponentials start, and the NEXT step's scores are loaded into
      //    the freed registers right there: wait S(j+1) (issued behind PV(j-1), long done), two tcgen05.ld, and the
      //    ~250 cycles of their latency pass under the chunk's 32 MUFUs instead of at the top of the next iteration;
      //  * as in v4: no per-step wait for PV(j-2), one p_full arrival per warp.
      if (nk <= 2) release_q();   // short items: the producer needs the buffer back sooner (two items ahead)
      float m_run = -INFINITY;  // running reference maximum, in log2 units (score * scale * log2 e)
      float l_run = 0.0f;
      uint32_t va[32], vb[32];
      {  // the item's first scores
        const uint32_t sbuf0 = paired ? (g >> 1) & 1u : g & 1u;
        if (wg_tid == 0) ATRACE(2 + wg, g, 0);
        mbar_wait(b_sfull + (wg * 2 + sbuf0) * 8, (sbuf0 ? scnt1 : scnt0) & 1u);
        if (sbuf0) ++scnt1; else ++scnt0;
        tc_fence_after_sync();
        if (wg_tid == 0) ATRACE(2 + wg, g, 1);
        tmem_ld32(t_tile + sbuf0 * A3_BKV, va);
        if (p.Tk > 32) tmem_ld32(t_tile + sbuf0 * A3_BKV + 32, vb);
      }
      for (int j = 0; j < nk; ++j) {
        // flat step of this tile's j-th key step: g + j, or g + 2 j + wg in paired-head mode (g is even there)
        const uint32_t sbuf = paired ? ((g >> 1) + static_cast<uint32_t>(j)) & 1u : (g + static_cast<uint32_t>(j)) & 1u;
        const uint32_t t_s = t_tile + sbuf * A3_BKV;
        const int rem = p.Tk - j * A3_BKV;
        const bool two = rem > 32;                   // second 32-key chunk holds a valid key
        const bool masked = flags[j] != 0;           // warp-uniform
        const float* cap_j = caps + j * A3_BKV;
        tmem_ld_wait();
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 2);
        if (masked) {
          apply_caps(va, cap_j);
          if (two) apply_caps(vb, cap_j + 32);
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        max4(m4, va);
        if (two) max4(m4, vb);
        const float tile_max = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;  // sc > 0
        // ---- lazy running maximum; O / l rescale only when the maximum grew by more than 2^TAU (rare)
        if (j == 0) {
          m_run = tile_max;
        } else {
          const bool need = tile_max > m_run + A3_LAZY_TAU;
          if (__any_sync(0xffffffffu, need)) {
            // every PV of this tile must have retired: the latest one is PV(pv_issued - 1); the one before retired
            // before this step's S landed (S(j) is issued behind PV(j-2)), so this wait can not see a stale phase
            const uint32_t kk = pv_issued - 1u;
            mbar_wait(b_pvdone + (wg * 2 + (kk & 1u)) * 8, (kk >> 1) & 1u);
            tc_fence_after_sync();
            const float alpha = need ? ex2_approx(m_run - tile_max) : 1.0f;
            if (need) m_run = tile_max;
            l_run *= alpha;
#pragma unroll 1
            for (int c = 0; c < DH / 32; ++c) {
              uint32_t v[32];
              tmem_ld32(t_o + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st32(t_o + c * 32, v);
            }
          }
        }
        const float neg_m = (m_run == -INFINITY) ? 0.0f : -m_run;
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 3);

        // ---- p = 2^(s*scale - m), row sum, bf16 P into the first 32 columns of this S buffer; the two warpgroups
        // take turns (named barriers 3 / 4) so that their exponentials do not collide on the quarter-rate XU pipe
#ifndef HRIEMO_ATTN_NO_PINGPONG
        asm volatile("bar.sync %0, 256;" ::"r"(3 + wg) : "memory");
#endif
        float l4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        uint32_t pk[16];
        float2 xs[16];
        scale32(va, sc, neg_m, xs);
        exp_pack_x(xs, l4, pk);
        tmem_st16(t_s, pk);
        if (two) scale32(vb, sc, neg_m, xs);     // vb is dead from here on
        // ---- the next step's scores, into the registers just freed.  S(j+1) is issued behind PV(j-1) and shares the tensor
        // pipe with the other tile, so it is often NOT there yet in the middle of this step (a blocking wait here, inside
        // the exp turn, cost 6 %): one non-blocking test now -- if it has landed its load latency passes under the
        // exponentials below -- and the blocking wait after them otherwise
        const bool more = j + 1 < nk;
        const uint32_t nbuf = sbuf ^ 1u;
        bool fetched = false;
        auto fetch_next = [&]() {
          if (nbuf) ++scnt1; else ++scnt0;
          tc_fence_after_sync();
          tmem_ld32(t_tile + nbuf * A3_BKV, va);
          if (p.Tk - (j + 1) * A3_BKV > 32) tmem_ld32(t_tile + nbuf * A3_BKV + 32, vb);
          fetched = true;
        };
        if (more) {
          const bool ok = mbar_test_wait(b_sfull + (wg * 2 + nbuf) * 8, (nbuf ? scnt1 : scnt0) & 1u);
          if (__all_sync(0xffffffffu, ok)) fetch_next();
        }
        if (two) {
          exp_pack_x(xs, l4, pk);
          tmem_st16(t_s + 16, pk);
        }
#ifndef HRIEMO_ATTN_NO_PINGPONG
        asm volatile("bar.arrive %0, 256;" ::"r"(4 - wg) : "memory");   // the other warpgroup's turn
#endif
        l_run += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 4);
        if (more && !fetched) {
          mbar_wait(b_sfull + (wg * 2 + nbuf) * 8, (nbuf ? scnt1 : scnt0) & 1u);
          fetch_next();
        }
        tmem_st_wait();
        tc_fence_before_sync();
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 5);
        __syncwarp();
        if (lane == 0) mbar_arrive(b_pfull + (wg * 2 + (pv_issued & 1u)) * 8);
        ++pv_issued;
        if (j == 0 && nk > 2) release_q();
        if (wg_tid == 0) ATRACE(2 + wg, g + j, 6);
      }
#endif

      // ---- epilogue: O / l -> bf16, staged in this item's (now dead) Q buffer and written with one
      // TMA store per warp; the 3-D tensor map (column, t, utterance) clips rows t >= Tq.
#ifdef HRIEMO_ATTN_V3_SOFTMAX
      consume_pv(pv_issued);
#else
      if constexpr (PSEP) {
        consume_pv(pv_issued);
      } else {
        const uint32_t kk = pv_issued - 1u;   // the item's last PV (its predecessor retired before the last S landed)
        mbar_wait(b_pvdone + (wg * 2 + (kk & 1u)) * 8, (kk >> 1) & 1u);
      }
#endif
      tc_fence_after_sync();
      // l == 0 (every key masked) -> inf -> NaN like torch.softmax; DROP: the kept probabilities' 1 / (1 - p)
      const float inv_l = DROP ? (1.0f / l_run) * p.drop_scale : 1.0f / l_run;
      // what a backward pass needs to rebuild P: ln sum_k exp(s_k * scale); this thread's row is its TMEM lane
      if (p.lse != nullptr && q0 + quad * 32 + lane < p.Tq)
        p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Tq + q0 + quad * 32 + lane] = (m_run + log2f(l_run)) * 0.6931471805599453f;
      const uint32_t qslot = wg * 2 + (qcnt_w & 1u);
      const uint32_t stage_warp = sQ + qslot * L::Q_TILE + static_cast<uint32_t>(quad) * (32 * DH * 2);
      const uint32_t stage_row = stage_warp + static_cast<uint32_t>(lane) * (DH * 2);
      // all DH/32 loads are in flight before the first wait (the S registers are dead here): one
      // tensor-memory round trip per item instead of DH/32
      constexpr int OB = (DH / 32 > 3) ? 2 : DH / 32;   // chunks per batch (dh = 128 would spill with 4)
#pragma unroll
      for (int c0 = 0; c0 < DH / 32; c0 += OB) {
        uint32_t vo[OB][32];
#pragma unroll
        for (int c = 0; c < OB; ++c) tmem_ld32(t_o + (c0 + c) * 32, vo[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < OB; ++c) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + (c0 + c) * 64 + g4 * 16),
                         "r"(pack_bf16(__uint_as_float(vo[c][g4 * 8 + 0]) * inv_l, __uint_as_float(vo[c][g4 * 8 + 1]) * inv_l)),
                         "r"(pack_bf16(__uint_as_float(vo[c][g4 * 8 + 2]) * inv_l, __uint_as_float(vo[c][g4 * 8 + 3]) * inv_l)),
                         "r"(pack_bf16(__uint_as_float(vo[c][g4 * 8 + 4]) * inv_l, __uint_as_float(vo[c][g4 * 8 + 5]) * inv_l)),
                         "r"(pack_bf16(__uint_as_float(vo[c][g4 * 8 + 6]) * inv_l, __uint_as_float(vo[c][g4 * 8 + 7]) * inv_l))
                         : "memory");
          }
        }
      }
      tc_fence_before_sync();   // O reads retire before the next item's first p_full lets PV overwrite O
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (q0 + quad * 32 < p.Tq) {
          tma_store_3d(&tm_o, stage_warp, h * DH, q0 + quad * 32, b);
          bulk_commit();
        }
      }
      // The Q buffer (now the store's source) is handed back to the producer one step LATER, after the
      // next item's first hand-off, when the store has long since read it: waiting here would sit on
      // the critical path of every item (1 000 - 3 000 cycles; the trace of the 500 x 64 shape).
      pending_qslot = static_cast<int>(qslot);
      if (wg_tid == 0) ATRACE(2 + wg, g + nk - 1, 7);
      ++qcnt_w;
    }
    release_q();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


// =====================================================================================================================
// attention_fwd4_kernel: the long-key general form (T_q > 128, more than two key steps) with FOUR softmax warpgroups that
// split every 64-key step of a tile by COLUMNS.
//
// What the profile of the form above says (profiles/r02_attention_issue_notes.txt items 7-9): of the ~2 000 cycles a
// softmax warp spends per 64-key step only ~690 are exponentials; ~850 are the latencies of the synchronising
// instructions themselves -- mbarrier try_wait (S full), tcgen05.wait::ld, tcgen05.wait::st + fence, syncwarp + mbarrier
// arrive -- which cost their ~200 cycles each whether or not the data arrived long ago, block the warp, and with ONE
// softmax warp per scheduler and tile have nothing to overlap with.  More steps in flight per warp do not help (the waits
// block it); more warps per scheduler do.  Splitting a tile's key STEPS between two warpgroups lost the S look-ahead
// (item 5).  Here the two warpgroups of a tile work on the SAME step: warpgroup `half` owns keys [32 half, 32 half + 32)
// of every step, each thread still owns one query row (its TMEM lane), so S stays double-buffered and issued two steps
// ahead exactly as above, every warp runs the same chain on half the scores, and four warps per scheduler overlap each
// other's waits.  The price: the row maximum of a step spans both warpgroups -- the two threads of a row exchange their
// partial maxima through shared memory behind a 64-thread named barrier (one per tile and lane quadrant) before the
// exponentials, which also orders the P stores (P overlays S columns [0, 32): half 1's share lands on columns half 0 has
// to have read first).  Row sums are partial per thread and added in the epilogue; each thread scales and stages half of
// its row's O columns.  640 threads: warp 0 TMA, warps 1 / 2 the MMA issuers of tiles 0 / 1, warp 3 idle, warps 4..19
// softmax (tile = (w - 4) >> 3, half = ((w - 4) >> 2) & 1, quadrant = w & 3).  TMA producer and issuers are the general
// non-DEFER code of the kernel above.
constexpr int A4_THREADS = 640;
constexpr bool kFwd4Default = false;   // flipped once the form is measured faster on the B200

template <int DH>
struct Attn4Smem : Attn3Smem<DH> {
  // caps[2][n_kv*64] f32, flags[2][n_kv] i32 (one copy per TILE, built by its two warpgroups together), xm[2][2][2][128]
  // f32 (partial row maxima, double-buffered by step parity), xl[2][2][128] f32 (partial row sums)
  static int dyn_bytes4(int n_kv) {
    return Attn3Smem<DH>::DYN_OFF + 2 * n_kv * A3_BKV * 4 + 2 * n_kv * 4 + (2 * 2 * 2 * 128 + 2 * 2 * 128) * 4 + 1024;
  }
};

template <int DH, bool DROP = false>
__global__ void __launch_bounds__(A4_THREADS, 1)
attention_fwd4_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                      const Attn3Params p) {
  using L = Attn4Smem<DH>;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t TILE_COLS = 256;   // per query tile: S0 at +0, S1 at +64, O at +128
  constexpr uint32_t O_COL = 128;
  constexpr int KS = L::KV_STAGES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t sQ = base + L::Q_OFF, sK = base + L::K_OFF, sV = base + L::V_OFF;
  const uint32_t bars = base + L::BAR_OFF;
  const uint32_t b_qfull = bars + 0 * 8, b_qempty = bars + 4 * 8, b_kfull = bars + 8 * 8, b_kempty = bars + 11 * 8;
  const uint32_t b_vfull = bars + 14 * 8, b_vempty = bars + 17 * 8, b_sfull = bars + 20 * 8, b_pfull = bars + 24 * 8;
  const uint32_t b_pvdone = bars + 28 * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::TMEM_SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = p.n_kv;
  const uint32_t n_items_all = static_cast<uint32_t>(p.n_items);
  const uint32_t per_cta = (n_items_all + gridDim.x - 1) / gridDim.x;
  const uint32_t item_first = p.contiguous ? blockIdx.x * per_cta : blockIdx.x;
  const uint32_t item_stride = p.contiguous ? 1u : gridDim.x;
  const uint32_t item_last = p.contiguous ? min(item_first + per_cta, n_items_all) : n_items_all;  // exclusive
  auto steps_of = [&](uint32_t item) -> int {
    if (p.kv_steps == nullptr) return n_kv;
    return __ldg(p.kv_steps + item / static_cast<uint32_t>(p.items_per_b));
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    for (int s = 0; s < 4; ++s) {
      mbar_init(b_qfull + s * 8, 1);
      mbar_init(b_qempty + s * 8, 4);   // released by the four half-0 warps of the tile (they issue the O stores)
      mbar_init(b_sfull + s * 8, 1);
      mbar_init(b_pvdone + s * 8, 1);
      mbar_init(b_pfull + s * 8, 8);    // one arrival per softmax warp of the tile (both halves)
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(b_kfull + s * 8, 1);
      mbar_init(b_kempty + s * 8, 2);   // one arrival per MMA issuer
      mbar_init(b_vfull + s * 8, 1);
      mbar_init(b_vempty + s * 8, 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(base + L::TMEM_SLOT_OFF);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (general mode of attention_fwd3_kernel) =====================
    if (elect_one()) {
      uint32_t qcnt[2] = {0, 0};
      uint32_t g = 0;
      for (uint32_t item = item_first; item < item_last; item += item_stride) {
        const uint32_t bh = item / static_cast<uint32_t>(p.n_qp);
        const int qp = static_cast<int>(item - bh * static_cast<uint32_t>(p.n_qp));
        const int b = static_cast<int>(bh / static_cast<uint32_t>(p.H));
        const int h = static_cast<int>(bh - static_cast<uint32_t>(b) * static_cast<uint32_t>(p.H));
        const int q0 = qp * 2 * A3_BQ;
        for (int t = 0; t < 2; ++t) {
          if (q0 + t * A3_BQ >= p.Tq) break;
          const uint32_t qb = qcnt[t] & 1u, qpar = (qcnt[t] >> 1) & 1u;
          const uint32_t bar = (t * 2 + qb) * 8;
          mbar_wait(b_qempty + bar, qpar ^ 1);
          mbar_arrive_expect_tx(b_qfull + bar, L::Q_TILE);
          for (int c = 0; c < L::QCH; ++c)
            tma_load_2d(&tm_q, b_qfull + bar, sQ + (t * 2 + qb) * L::Q_TILE + c * L::Q_CHUNK, h * DH + c * 64,
                        b * p.Tq + q0 + t * A3_BQ);
          ++qcnt[t];
        }
        const int nk = steps_of(item);
        for (int j = 0; j < nk; ++j, ++g) {
          const uint32_t s = g % KS, par = (g / KS) & 1u;
          mbar_wait(b_kempty + s * 8, par ^ 1);
          mbar_arrive_expect_tx(b_kfull + s * 8, L::K_STAGE);
          for (int c = 0; c < L::QCH; ++c)
            tma_load_2d(&tm_k, b_kfull + s * 8, sK + s * L::K_STAGE + c * L::K_CHUNK, h * DH + c * 64,
                        b * p.Tk + j * A3_BKV);
          mbar_wait(b_vempty + s * 8, par ^ 1);
          mbar_arrive_expect_tx(b_vfull + s * 8, L::V_STAGE);
          for (int c = 0; c < DH / 32; ++c)
            tma_load_3d(&tm_v, b_vfull + s * 8, sV + s * L::V_STAGE + c * L::V_GROUP, h * DH + c * 32, j * A3_BKV, b);
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: warp 1 drives query tile 0, warp 2 query tile 1 (general non-DEFER form) =====
    auto run_issuer = [&](auto tile_c) {
      constexpr int t = decltype(tile_c)::value;
      if (tmem_base != 0u) __trap();   // this CTA owns all 512 columns: the base is the constant 0 (uniform descriptors)
      constexpr uint32_t idesc_s = umma_idesc_bf16(A3_BQ, A3_BKV);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(A3_BQ, DH) | kUmmaBMajorMN;
      const uint32_t tile_tmem = t * TILE_COLS;
      const uint64_t k_desc0 = umma_desc_sw128(sK);
      const uint64_t v_desc0 = umma_desc_mn_sw64(sV, L::V_GROUP);
      auto tile_active = [&](uint32_t item) {
        return t == 0 || static_cast<int>(item % static_cast<uint32_t>(p.n_qp)) * 2 * A3_BQ + A3_BQ < p.Tq;
      };
      uint32_t s_g = 0, s_item = item_first, qcnt = 0;
      int s_f = 0;
      int s_nk = s_item < item_last ? steps_of(s_item) : 0;
      bool s_act = tile_active(s_item);
      auto issue_s = [&]() {
        const uint32_t ks = s_g % KS, kpar = (s_g / KS) & 1u;
        const uint32_t qslot = t * 2 + (qcnt & 1u);
        if (s_act && s_f == 0) mbar_wait(b_qfull + qslot * 8, (qcnt >> 1) & 1u);
        mbar_wait(b_kfull + ks * 8, kpar);
        if (s_act) {
          tc_fence_after_sync();
          const uint64_t q_desc = umma_desc_sw128(sQ + qslot * L::Q_TILE);
          const uint64_t k_desc = k_desc0 + ((ks * L::K_STAGE) >> 4);
          const uint32_t d_tmem = tile_tmem + (s_g & 1u) * A3_BKV;
#pragma unroll
          for (int st = 0; st < DH / 16; ++st) {
            umma_bf16(d_tmem, q_desc + (((st >> 2) * L::Q_CHUNK + (st & 3) * 32) >> 4),
                      k_desc + (((st >> 2) * L::K_CHUNK + (st & 3) * 32) >> 4), idesc_s, st != 0);
          }
          umma_commit(b_sfull + (t * 2 + (s_g & 1u)) * 8);
          umma_commit(b_kempty + ks * 8);
          if (s_f == s_nk - 1) ++qcnt;
        } else {
          mbar_arrive(b_kempty + ks * 8);
        }
        ++s_g;
        if (++s_f == s_nk) {
          s_f = 0;
          s_item += item_stride;
          s_act = s_item < item_last && tile_active(s_item);
          s_nk = s_item < item_last ? steps_of(s_item) : 0;
        }
      };
      if (s_item < item_last) issue_s();
      if (s_item < item_last) issue_s();
      uint32_t pcnt = 0, pv_item = item_first;
      int pv_f = 0;
      int pv_nk = pv_item < item_last ? steps_of(pv_item) : 0;
      bool pv_act = tile_active(pv_item);
      for (uint32_t g = 0; pv_item < item_last; ++g) {
        const uint32_t vs = g % KS, vpar = (g / KS) & 1u;
        mbar_wait(b_vfull + vs * 8, vpar);
        if (pv_act) {
          const int rem = p.Tk - pv_f * A3_BKV;  // keys left from this step on (> 0)
          const uint32_t slot = t * 2 + (pcnt & 1u);
          mbar_wait(b_pfull + slot * 8, (pcnt >> 1) & 1u);
          ++pcnt;
          tc_fence_after_sync();
          const uint64_t v_desc = v_desc0 + ((vs * L::V_STAGE) >> 4);
          const uint32_t p_tmem = tile_tmem + (g & 1u) * A3_BKV;
#pragma unroll
          for (int st = 0; st < A3_BKV / 16; ++st) {
            if (st * 16 < rem)
              umma_bf16_ts(tile_tmem + O_COL, p_tmem + st * 8, v_desc + ((st * 16 * 64) >> 4), idesc_pv, (pv_f | st) != 0);
          }
          umma_commit(b_pvdone + slot * 8);
          umma_commit(b_vempty + vs * 8);
        } else {
          mbar_arrive(b_vempty + vs * 8);
        }
        if (s_item < item_last) issue_s();  // S(g+2) reuses the S buffer whose P was just consumed
        if (++pv_f == pv_nk) {
          pv_f = 0;
          pv_item += item_stride;
          pv_act = pv_item < item_last && tile_active(pv_item);
          pv_nk = pv_item < item_last ? steps_of(pv_item) : 0;
        }
      }
    };
    if (elect_one()) {
      if (warp == 1) run_issuer(std::integral_constant<int, 0>{});
      else run_issuer(std::integral_constant<int, 1>{});
    }
  } else if (warp >= 4) {
    // ===================== softmax: four warpgroups, (tile, half) = (0,0) (0,1) (1,0) (1,1) =====================
    const int swg = (warp - 4) >> 2;
    const int tile = swg >> 1, half = swg & 1;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                       // query row inside the tile = TMEM lane
    const int tile_tid = threadIdx.x - 128 - tile * 256;    // thread within the tile's two warpgroups
    float* dyn = reinterpret_cast<float*>(base_ptr + L::DYN_OFF);
    float* caps = dyn + tile * n_kv * A3_BKV;
    int* flags = reinterpret_cast<int*>(dyn + 2 * n_kv * A3_BKV) + tile * n_kv;
    float* xm = dyn + 2 * n_kv * A3_BKV + 2 * n_kv;         // [2][2][2][128]
    float* xl = xm + 2 * 2 * 2 * 128;                       // [2][2][128]
    const uint32_t t_tile = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + tile * TILE_COLS;
    const uint32_t t_o = t_tile + O_COL;
    const uint32_t pid = 5 + tile * 4 + quad;               // named barrier of the two warps that share this quadrant of the tile
    const float sc = p.scale_log2;
    constexpr int OH = DH / 2;                              // O columns this thread rescales / stages
    uint32_t scnt0 = 0, scnt1 = 0, pv_issued = 0, xstep = 0;
    uint32_t g = 0;
    int pending_qslot = -1;
    auto release_q = [&]() {   // half-0 warps only (they issue the O stores)
      if (pending_qslot >= 0) {
        if (lane == 0) {
          bulk_wait_read<0>();
          mbar_arrive(b_qempty + pending_qslot * 8);
        }
        pending_qslot = -1;
      }
    };
    uint32_t qcnt_w = 0;
    int cur_b = -1;
    int nk = 0;
    for (uint32_t item = item_first; item < item_last; item += item_stride, g += nk) {
      nk = steps_of(item);
      const int b = static_cast<int>(item / static_cast<uint32_t>(p.items_per_b));
      const uint32_t in_b = item - static_cast<uint32_t>(b) * static_cast<uint32_t>(p.items_per_b);
      const int h = static_cast<int>(in_b / static_cast<uint32_t>(p.n_qp));
      const int q0 = static_cast<int>(in_b % static_cast<uint32_t>(p.n_qp)) * 2 * A3_BQ + tile * A3_BQ;
      if (q0 >= p.Tq) continue;   // this tile does not exist for the item
      [[maybe_unused]] const uint32_t dbase =
          drop_key_bh(p.drop_key, static_cast<uint32_t>(b * p.H + h)) + static_cast<uint32_t>(q0 + row) * DROP_C_ROW;
      if (cur_b < 0 || (b != cur_b && p.key_pad != nullptr)) {   // without a mask the caps only depend on Tk
        asm volatile("bar.sync %0, 256;" ::"r"(1 + tile) : "memory");   // both warpgroups of the tile are past the old caps
        for (int j = tile_tid; j < n_kv; j += 256) flags[j] = 0;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + tile) : "memory");
        for (int kk = tile_tid; kk < n_kv * A3_BKV; kk += 256) {
          bool pad = kk >= p.Tk;
          if (!pad && p.key_pad != nullptr) pad = p.key_pad[static_cast<int64_t>(b) * p.Tk + kk] != 0;
          caps[kk] = pad ? -INFINITY : INFINITY;
          if (pad) flags[kk / A3_BKV] = 1;
        }
        asm volatile("bar.sync %0, 256;" ::"r"(1 + tile) : "memory");
        cur_b = b;
      }
      if (half == 0 && nk <= 2) release_q();
      float m_run = -INFINITY;   // reference maximum of the row, identical in the two threads that share it
      float l_run = 0.0f;        // this thread's share of the row sum
      uint32_t v[32];
      {  // the item's first scores: this warpgroup's 32 columns of the S buffer
        const uint32_t sbuf0 = g & 1u;
        mbar_wait(b_sfull + (tile * 2 + sbuf0) * 8, (sbuf0 ? scnt1 : scnt0) & 1u);
        if (sbuf0) ++scnt1; else ++scnt0;
        tc_fence_after_sync();
        tmem_ld32(t_tile + sbuf0 * A3_BKV + half * 32, v);
      }
      for (int j = 0; j < nk; ++j) {
        const uint32_t sbuf = (g + static_cast<uint32_t>(j)) & 1u;
        const uint32_t t_s = t_tile + sbuf * A3_BKV;
        const bool masked = flags[j] != 0;           // warp-uniform
        tmem_ld_wait();
        if (masked) apply_caps(v, caps + j * A3_BKV + half * 32);
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        max4(m4, v);
        const float own_max = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;   // sc > 0
        // ---- the row maximum spans both halves: exchange, and (the same barrier) both halves hold their scores in
        // registers, so P may overwrite S columns [0, 32)
        const uint32_t xb = xstep & 1u;
        ++xstep;
        xm[((xb * 2 + tile) * 2 + half) * 128 + row] = own_max;
        asm volatile("bar.sync %0, 64;" ::"r"(pid) : "memory");
        const float tile_max = fmaxf(own_max, xm[((xb * 2 + tile) * 2 + (half ^ 1)) * 128 + row]);
        if (j == 0) {
          m_run = tile_max;
        } else {
          const bool need = tile_max > m_run + A3_LAZY_TAU;
          if (__any_sync(0xffffffffu, need)) {   // the partner warp holds the same rows: it takes the same branch
            const uint32_t kk = pv_issued - 1u;  // every PV of this tile must have retired (PV(j-2) did before S(j) landed)
            mbar_wait(b_pvdone + (tile * 2 + (kk & 1u)) * 8, (kk >> 1) & 1u);
            tc_fence_after_sync();
            const float alpha = need ? ex2_approx(m_run - tile_max) : 1.0f;
            if (need) m_run = tile_max;
            l_run *= alpha;
#pragma unroll 1
            for (int c = 0; c < OH / 16; ++c) {   // this thread's half of the row's O columns
              uint32_t vo[16];
              tmem_ld16(t_o + half * OH + c * 16, vo);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) vo[i] = __float_as_uint(__uint_as_float(vo[i]) * alpha);
              tmem_st16(t_o + half * OH + c * 16, vo);
            }
          }
        }
        const float neg_m = (m_run == -INFINITY) ? 0.0f : -m_run;
        // ---- p = 2^(s*scale - m) for this thread's 32 keys; bf16 P into columns [16 half, 16 half + 16) of the S buffer
        float l4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        uint32_t pk[16];
        {
          float2 xs[16];
          scale32(v, sc, neg_m, xs);
          exp_pack_x(xs, l4, pk);
        }
        if constexpr (DROP) drop_pk(pk, dbase, static_cast<uint32_t>(j) * 16u + static_cast<uint32_t>(half) * 8u, p.drop_p8);
        tmem_st16(t_s + half * 16, pk);
        l_run += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        // ---- the next step's scores while the P stores drain
        if (j + 1 < nk) {
          const uint32_t nbuf = sbuf ^ 1u;
          mbar_wait(b_sfull + (tile * 2 + nbuf) * 8, (nbuf ? scnt1 : scnt0) & 1u);
          if (nbuf) ++scnt1; else ++scnt0;
          tc_fence_after_sync();
          tmem_ld32(t_tile + nbuf * A3_BKV + half * 32, v);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_pfull + (tile * 2 + (pv_issued & 1u)) * 8);
        ++pv_issued;
        if (half == 0 && j == 0 && nk > 2) release_q();
      }
      // ---- epilogue: this thread's half of O / l -> bf16, staged in the item's (now dead) Q buffer; the half-0 warp of a
      // quadrant issues the TMA store of the 32-row box once both halves have written
      {
        const uint32_t kk = pv_issued - 1u;
        mbar_wait(b_pvdone + (tile * 2 + (kk & 1u)) * 8, (kk >> 1) & 1u);
      }
      tc_fence_after_sync();
      xl[(tile * 2 + half) * 128 + row] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(pid) : "memory");
      const float l_tot = xl[(tile * 2 + 0) * 128 + row] + xl[(tile * 2 + 1) * 128 + row];   // fixed order in both threads
      const float inv_l = DROP ? (1.0f / l_tot) * p.drop_scale : 1.0f / l_tot;   // l == 0 -> NaN like torch.softmax
      if (half == 0 && p.lse != nullptr && q0 + row < p.Tq)
        p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Tq + q0 + row] = (m_run + log2f(l_tot)) * 0.6931471805599453f;
      const uint32_t qslot = tile * 2 + (qcnt_w & 1u);
      const uint32_t stage_warp = sQ + qslot * L::Q_TILE + static_cast<uint32_t>(quad) * (32 * DH * 2);
      const uint32_t stage_row = stage_warp + static_cast<uint32_t>(lane) * (DH * 2) + static_cast<uint32_t>(half) * (OH * 2);
#pragma unroll
      for (int c = 0; c < OH / 16; ++c) {
        uint32_t vo[16];
        tmem_ld16(t_o + half * OH + c * 16, vo);
        tmem_ld_wait();
#pragma unroll
        for (int g4 = 0; g4 < 2; ++g4) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + c * 32 + g4 * 16),
                       "r"(pack_bf16(__uint_as_float(vo[g4 * 8 + 0]) * inv_l, __uint_as_float(vo[g4 * 8 + 1]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(vo[g4 * 8 + 2]) * inv_l, __uint_as_float(vo[g4 * 8 + 3]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(vo[g4 * 8 + 4]) * inv_l, __uint_as_float(vo[g4 * 8 + 5]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(vo[g4 * 8 + 6]) * inv_l, __uint_as_float(vo[g4 * 8 + 7]) * inv_l))
                       : "memory");
        }
      }
      tc_fence_before_sync();   // O reads retire before the next item's first p_full lets PV overwrite O
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 64;" ::"r"(pid) : "memory");   // both halves of the 32-row box are staged (and xl is read)
      if (half == 0) {
        if (lane == 0 && q0 + quad * 32 < p.Tq) {
          tma_store_3d(&tm_o, stage_warp, h * DH, q0 + quad * 32, b);
          bulk_commit();
        }
        pending_qslot = static_cast<int>(qslot);
      }
      ++qcnt_w;
    }
    if (half == 0) release_q();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


// =====================================================================================================================
// attention_fwd5_kernel: the long-key general form with the O / l epilogue on its OWN warpgroup.
//
// What the warp-state samples of the form at the top say (profiles/r02_ncu_attention_v16_summary.txt): a softmax warp
// spends 22 % of its time per work item NOT producing probabilities -- 6.5 % waiting for the item's last PV, 11 % scaling
// O by 1 / l, packing it and staging it for the TMA store (with 4-way bank conflicts), 4.5 % setting up the next item --
// and for all of that time its query tile issues no tensor work beyond the two S tiles of the look-ahead.  Here the
// softmax warpgroup of a tile ends an item by handing its row sums (l, m) to a fourth warpgroup through shared memory
// and goes straight on to the next item's first scores (already in tensor memory); the epilogue warpgroup waits for the
// last PV, pulls O into registers, tells the tile's MMA issuer that O may be overwritten (o_free: the next item's first
// PV does not accumulate), and only then scales / packs / stages / stores.  By the time the next item's first P is
// handed over (a whole softmax step) O has long been read.
//
// 512 threads = 4 warpgroups (register budget moved between them with setmaxnreg: 56 / 168 / 168 / 120 per thread):
//   warps 0..3   TMA producer, MMA issuer of tile 0, MMA issuer of tile 1, idle
//   warps 4..7   softmax of query tile 0 \ the v4 loop of attention_fwd3_kernel, one query row per thread
//   warps 8..11  softmax of query tile 1 /
//   warps 12..15 epilogues of both tiles, one query row per thread (TMEM lane quadrant = warp & 3)
// TMA producer and issuers are the general non-DEFER code of attention_fwd3_kernel.
//
// Measured (tools/attn_sweep.py, same build, same box, T_q = T_k = T, dh = 96): correct (the attention and dropout tests pass
// with HRIEMO_ATTN_FWD5=1, bit-identical to the shipped form) and 2 - 7 % FASTER for short items (T = 200 .. 384, i.e. up to
// six key steps: 300 x 300 0.96 - 1.00 -> 0.92 - 0.94 ms at B = 1408), equal at T = 448, 3 - 7 % SLOWER from T = 500 up
// (500 x 500 0.510 - 0.518 -> 0.524 - 0.539 ms): the per-item overhead is gone, but the per-step loop of the softmax warps
// is slower with four more warps on their schedulers (handing (l, m) over through a named barrier instead of an mbarrier
// the epilogue warps poll changed nothing measurable).  Not the default; kept behind HRIEMO_ATTN_FWD5=1.
constexpr int A5_THREADS = 512;
constexpr bool kFwd5Default = false;   // flipped once the form is measured faster on the B200

template <int DH>
struct Attn5Smem : Attn3Smem<DH> {
  // behind the barriers of the base layout: epi_full[2] epi_empty[2] o_free[2] (64 bytes reserved), then caps[2][n_kv*64]
  // f32, flags[2][n_kv] i32, xl[2][128] f32 (row sums), xm[2][128] f32 (reference maxima, log2 units)
  static constexpr int BAR5_OFF = Attn3Smem<DH>::DYN_OFF;
  static constexpr int DYN5_OFF = BAR5_OFF + 64;
  static int dyn_bytes5(int n_kv) { return DYN5_OFF + 2 * n_kv * A3_BKV * 4 + 2 * n_kv * 4 + 2 * 2 * 128 * 4 + 1024; }
};

template <int DH, bool DROP = false>
__global__ void __launch_bounds__(A5_THREADS, 1)
attention_fwd5_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                      const Attn3Params p) {
  using L = Attn5Smem<DH>;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t TILE_COLS = 256;   // per query tile: S0 at +0, S1 at +64, O at +128
  constexpr uint32_t O_COL = 128;
  constexpr int KS = L::KV_STAGES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t sQ = base + L::Q_OFF, sK = base + L::K_OFF, sV = base + L::V_OFF;
  const uint32_t bars = base + L::BAR_OFF;
  const uint32_t b_qfull = bars + 0 * 8, b_qempty = bars + 4 * 8, b_kfull = bars + 8 * 8, b_kempty = bars + 11 * 8;
  const uint32_t b_vfull = bars + 14 * 8, b_vempty = bars + 17 * 8, b_sfull = bars + 20 * 8, b_pfull = bars + 24 * 8;
  const uint32_t b_pvdone = bars + 28 * 8;
  const uint32_t b_epifull = base + L::BAR5_OFF;          // [tile] the item's (l, m) are in shared memory
  const uint32_t b_epiempty = base + L::BAR5_OFF + 16;    // [tile] ... and have been read
  const uint32_t b_ofree = base + L::BAR5_OFF + 32;       // [tile] the item's O is in the epilogue warps' registers
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::TMEM_SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = p.n_kv;
  const uint32_t n_items_all = static_cast<uint32_t>(p.n_items);
  const uint32_t per_cta = (n_items_all + gridDim.x - 1) / gridDim.x;
  const uint32_t item_first = p.contiguous ? blockIdx.x * per_cta : blockIdx.x;
  const uint32_t item_stride = p.contiguous ? 1u : gridDim.x;
  const uint32_t item_last = p.contiguous ? min(item_first + per_cta, n_items_all) : n_items_all;  // exclusive
  auto steps_of = [&](uint32_t item) -> int {
    if (p.kv_steps == nullptr) return n_kv;
    return __ldg(p.kv_steps + item / static_cast<uint32_t>(p.items_per_b));
  };
  float* dyn = reinterpret_cast<float*>(base_ptr + L::DYN5_OFF);
  float* xl = dyn + 2 * n_kv * A3_BKV + 2 * n_kv;   // [2][128]
  float* xm = xl + 2 * 128;                         // [2][128]

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    for (int s = 0; s < 4; ++s) {
      mbar_init(b_qfull + s * 8, 1);
      mbar_init(b_qempty + s * 8, 4);   // released by the four epilogue warps (each stores its 32 rows)
      mbar_init(b_sfull + s * 8, 1);
      mbar_init(b_pvdone + s * 8, 1);
      mbar_init(b_pfull + s * 8, 4);    // one arrival per softmax warp of the tile
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(b_kfull + s * 8, 1);
      mbar_init(b_kempty + s * 8, 2);   // one arrival per MMA issuer
      mbar_init(b_vfull + s * 8, 1);
      mbar_init(b_vempty + s * 8, 2);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_epifull + s * 8, 4);   // one arrival per softmax warp of the tile
      mbar_init(b_epiempty + s * 8, 4);  // one arrival per epilogue warp
      mbar_init(b_ofree + s * 8, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(base + L::TMEM_SLOT_OFF);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int wgrp = warp >> 2;

  if (wgrp == 0) {
    reg_dealloc<56>();
    if (warp == 0) {
      // ===================== TMA producer (general mode of attention_fwd3_kernel) =====================
      if (elect_one()) {
        uint32_t qcnt[2] = {0, 0};
        uint32_t g = 0;
        for (uint32_t item = item_first; item < item_last; item += item_stride) {
          const uint32_t bh = item / static_cast<uint32_t>(p.n_qp);
          const int qp = static_cast<int>(item - bh * static_cast<uint32_t>(p.n_qp));
          const int b = static_cast<int>(bh / static_cast<uint32_t>(p.H));
          const int h = static_cast<int>(bh - static_cast<uint32_t>(b) * static_cast<uint32_t>(p.H));
          const int q0 = qp * 2 * A3_BQ;
          for (int t = 0; t < 2; ++t) {
            if (q0 + t * A3_BQ >= p.Tq) break;
            const uint32_t qb = qcnt[t] & 1u, qpar = (qcnt[t] >> 1) & 1u;
            const uint32_t bar = (t * 2 + qb) * 8;
            mbar_wait(b_qempty + bar, qpar ^ 1);
            mbar_arrive_expect_tx(b_qfull + bar, L::Q_TILE);
            for (int c = 0; c < L::QCH; ++c)
              tma_load_2d(&tm_q, b_qfull + bar, sQ + (t * 2 + qb) * L::Q_TILE + c * L::Q_CHUNK, h * DH + c * 64,
                          b * p.Tq + q0 + t * A3_BQ);
            ++qcnt[t];
          }
          const int nk = steps_of(item);
          for (int j = 0; j < nk; ++j, ++g) {
            const uint32_t s = g % KS, par = (g / KS) & 1u;
            mbar_wait(b_kempty + s * 8, par ^ 1);
            mbar_arrive_expect_tx(b_kfull + s * 8, L::K_STAGE);
            for (int c = 0; c < L::QCH; ++c)
              tma_load_2d(&tm_k, b_kfull + s * 8, sK + s * L::K_STAGE + c * L::K_CHUNK, h * DH + c * 64,
                          b * p.Tk + j * A3_BKV);
            mbar_wait(b_vempty + s * 8, par ^ 1);
            mbar_arrive_expect_tx(b_vfull + s * 8, L::V_STAGE);
            for (int c = 0; c < DH / 32; ++c)
              tma_load_3d(&tm_v, b_vfull + s * 8, sV + s * L::V_STAGE + c * L::V_GROUP, h * DH + c * 32, j * A3_BKV, b);
          }
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuers: warp 1 drives query tile 0, warp 2 query tile 1 (general non-DEFER form) =====
      auto run_issuer = [&](auto tile_c) {
        constexpr int t = decltype(tile_c)::value;
        if (tmem_base != 0u) __trap();   // this CTA owns all 512 columns: the base is the constant 0 (uniform descriptors)
        constexpr uint32_t idesc_s = umma_idesc_bf16(A3_BQ, A3_BKV);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(A3_BQ, DH) | kUmmaBMajorMN;
        const uint32_t tile_tmem = t * TILE_COLS;
        const uint64_t k_desc0 = umma_desc_sw128(sK);
        const uint64_t v_desc0 = umma_desc_mn_sw64(sV, L::V_GROUP);
        auto tile_active = [&](uint32_t item) {
          return t == 0 || static_cast<int>(item % static_cast<uint32_t>(p.n_qp)) * 2 * A3_BQ + A3_BQ < p.Tq;
        };
        uint32_t s_g = 0, s_item = item_first, qcnt = 0;
        int s_f = 0;
        int s_nk = s_item < item_last ? steps_of(s_item) : 0;
        bool s_act = tile_active(s_item);
        auto issue_s = [&]() {
          const uint32_t ks = s_g % KS, kpar = (s_g / KS) & 1u;
          const uint32_t qslot = t * 2 + (qcnt & 1u);
          if (s_act && s_f == 0) mbar_wait(b_qfull + qslot * 8, (qcnt >> 1) & 1u);
          mbar_wait(b_kfull + ks * 8, kpar);
          if (s_act) {
            tc_fence_after_sync();
            const uint64_t q_desc = umma_desc_sw128(sQ + qslot * L::Q_TILE);
            const uint64_t k_desc = k_desc0 + ((ks * L::K_STAGE) >> 4);
            const uint32_t d_tmem = tile_tmem + (s_g & 1u) * A3_BKV;
#pragma unroll
            for (int st = 0; st < DH / 16; ++st) {
              umma_bf16(d_tmem, q_desc + (((st >> 2) * L::Q_CHUNK + (st & 3) * 32) >> 4),
                        k_desc + (((st >> 2) * L::K_CHUNK + (st & 3) * 32) >> 4), idesc_s, st != 0);
            }
            umma_commit(b_sfull + (t * 2 + (s_g & 1u)) * 8);
            umma_commit(b_kempty + ks * 8);
            if (s_f == s_nk - 1) ++qcnt;
          } else {
            mbar_arrive(b_kempty + ks * 8);
          }
          ++s_g;
          if (++s_f == s_nk) {
            s_f = 0;
            s_item += item_stride;
            s_act = s_item < item_last && tile_active(s_item);
            s_nk = s_item < item_last ? steps_of(s_item) : 0;
          }
        };
        if (s_item < item_last) issue_s();
        if (s_item < item_last) issue_s();
        uint32_t pcnt = 0, pv_item = item_first, ocnt = 0;   // ocnt: items of this tile whose first PV has been issued
        int pv_f = 0;
        int pv_nk = pv_item < item_last ? steps_of(pv_item) : 0;
        bool pv_act = tile_active(pv_item);
        for (uint32_t g = 0; pv_item < item_last; ++g) {
          const uint32_t vs = g % KS, vpar = (g / KS) & 1u;
          mbar_wait(b_vfull + vs * 8, vpar);
          if (pv_act) {
            const int rem = p.Tk - pv_f * A3_BKV;  // keys left from this step on (> 0)
            const uint32_t slot = t * 2 + (pcnt & 1u);
            mbar_wait(b_pfull + slot * 8, (pcnt >> 1) & 1u);
            ++pcnt;
            if (pv_f == 0) {
              // the item's first PV overwrites O: the previous item's O must be in the epilogue warps' registers
              if (ocnt > 0) mbar_wait(b_ofree + t * 8, (ocnt - 1u) & 1u);
              ++ocnt;
            }
            tc_fence_after_sync();
            const uint64_t v_desc = v_desc0 + ((vs * L::V_STAGE) >> 4);
            const uint32_t p_tmem = tile_tmem + (g & 1u) * A3_BKV;
#pragma unroll
            for (int st = 0; st < A3_BKV / 16; ++st) {
              if (st * 16 < rem)
                umma_bf16_ts(tile_tmem + O_COL, p_tmem + st * 8, v_desc + ((st * 16 * 64) >> 4), idesc_pv, (pv_f | st) != 0);
            }
            umma_commit(b_pvdone + slot * 8);
            umma_commit(b_vempty + vs * 8);
          } else {
            mbar_arrive(b_vempty + vs * 8);
          }
          if (s_item < item_last) issue_s();  // S(g+2) reuses the S buffer whose P was just consumed
          if (++pv_f == pv_nk) {
            pv_f = 0;
            pv_item += item_stride;
            pv_act = pv_item < item_last && tile_active(pv_item);
            pv_nk = pv_item < item_last ? steps_of(pv_item) : 0;
          }
        }
      };
      if (elect_one()) {
        if (warp == 1) run_issuer(std::integral_constant<int, 0>{});
        else run_issuer(std::integral_constant<int, 1>{});
      }
    }
  } else if (wgrp == 3) {
    // ===================== epilogue warpgroup: O / l of both tiles =====================
    reg_dealloc<120>();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    uint32_t done[2] = {0u, 0u};   // items finished per tile -> (l, m) phase, Q buffer
    uint32_t pvc[2] = {0u, 0u};    // PVs of the tile up to and including the current item
    for (uint32_t item = item_first; item < item_last; item += item_stride) {
      const int nk = steps_of(item);
      const int b = static_cast<int>(item / static_cast<uint32_t>(p.items_per_b));
      const uint32_t in_b = item - static_cast<uint32_t>(b) * static_cast<uint32_t>(p.items_per_b);
      const int h = static_cast<int>(in_b / static_cast<uint32_t>(p.n_qp));
      const int q_pair = static_cast<int>(in_b % static_cast<uint32_t>(p.n_qp)) * 2 * A3_BQ;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int q0 = q_pair + t * A3_BQ;
        if (q0 >= p.Tq) continue;
        const uint32_t t_o = tmem_base + lane_sel + t * TILE_COLS + O_COL;
        mbar_wait(b_epifull + t * 8, done[t] & 1u);
        const float l_run = xl[t * 128 + row], m_run = xm[t * 128 + row];
        __syncwarp();
        if (lane == 0) mbar_arrive(b_epiempty + t * 8);
        // the item's last PV (its predecessors retired before it: one issuer, in order).  Nothing can complete this
        // barrier's phase twice behind our back: the next PV that lands on it is the next item's SECOND one, and the
        // next item's first waits for the o_free below.
        pvc[t] += static_cast<uint32_t>(nk);
        const uint32_t kk = pvc[t] - 1u;
        mbar_wait(b_pvdone + (t * 2 + (kk & 1u)) * 8, (kk >> 1) & 1u);
        tc_fence_after_sync();
        // l == 0 (every key masked) -> inf -> NaN like torch.softmax; DROP: the kept probabilities' 1 / (1 - p)
        const float inv_l = DROP ? (1.0f / l_run) * p.drop_scale : 1.0f / l_run;
        if (p.lse != nullptr && q0 + row < p.Tq)
          p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Tq + q0 + row] = (m_run + log2f(l_run)) * 0.6931471805599453f;
        const uint32_t qslot = t * 2 + (done[t] & 1u);
        const uint32_t stage_warp = sQ + qslot * L::Q_TILE + static_cast<uint32_t>(quad) * (32 * DH * 2);
        const uint32_t stage_row = stage_warp + static_cast<uint32_t>(lane) * (DH * 2);
        constexpr int OB = (DH / 32 > 3) ? 2 : DH / 32;   // chunks per batch
#pragma unroll
        for (int c0 = 0; c0 < DH / 32; c0 += OB) {
          uint32_t vo[OB][32];
#pragma unroll
          for (int c = 0; c < OB; ++c) tmem_ld32(t_o + (c0 + c) * 32, vo[c]);
          tmem_ld_wait();
          if (c0 + OB >= DH / 32) {   // O is in registers: the tile's next item may start accumulating
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_ofree + t * 8);
          }
#pragma unroll
          for (int c = 0; c < OB; ++c) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + (c0 + c) * 64 + g4 * 16),
                           "r"(pack_bf16(__uint_as_float(vo[c][g4 * 8 + 0]) * inv_l, __uint_as_float(vo[c][g4 * 8 + 1]) * inv_l)),
                           "r"(pack_bf16(__uint_as_float(vo[c][g4 * 8 + 2]) * inv_l, __uint_as_float(vo[c][g4 * 8 + 3]) * inv_l)),
                           "r"(pack_bf16(__uint_as_float(vo[c][g4 * 8 + 4]) * inv_l, __uint_as_float(vo[c][g4 * 8 + 5]) * inv_l)),
                           "r"(pack_bf16(__uint_as_float(vo[c][g4 * 8 + 6]) * inv_l, __uint_as_float(vo[c][g4 * 8 + 7]) * inv_l))
                           : "memory");
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (q0 + quad * 32 < p.Tq) {
            tma_store_3d(&tm_o, stage_warp, h * DH, q0 + quad * 32, b);
            bulk_commit();
          }
          // this warpgroup is off the critical path: wait for the store to have read the staging bytes and hand the
          // Q buffer straight back (the producer needs it for the item after next)
          bulk_wait_read<0>();
          mbar_arrive(b_qempty + qslot * 8);
        }
        __syncwarp();
        ++done[t];
      }
    }
  } else {
    // ===================== softmax warpgroups (the v4 loop of attention_fwd3_kernel, general form) =====================
    reg_alloc<168>();
    const int wg = wgrp - 1;                 // query tile handled by this warpgroup
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int wg_tid = threadIdx.x - 128 - wg * 128;
    const int row = quad * 32 + lane;
    float* caps = dyn + wg * n_kv * A3_BKV;
    int* flags = reinterpret_cast<int*>(dyn + 2 * n_kv * A3_BKV) + wg * n_kv;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_tile = tmem_base + lane_sel + wg * TILE_COLS;
    const uint32_t t_o = t_tile + O_COL;
    const float sc = p.scale_log2;
    uint32_t scnt0 = 0, scnt1 = 0;   // completed uses of each S buffer of this tile
    uint32_t pv_issued = 0;          // P tiles handed to the MMA warp so far
    uint32_t items_done = 0;         // items of this tile handed to the epilogue warpgroup
    uint32_t g = 0;                  // flat step index of this CTA at the start of the current item
    if (wg == 1) asm volatile("bar.arrive %0, 256;" ::"r"(3) : "memory");   // warpgroup 0 takes the first turn
    int cur_b = -1;
    int nk = 0;
    for (uint32_t item = item_first; item < item_last; item += item_stride, g += nk) {
      nk = steps_of(item);
      const int b = static_cast<int>(item / static_cast<uint32_t>(p.items_per_b));
      const uint32_t in_b = item - static_cast<uint32_t>(b) * static_cast<uint32_t>(p.items_per_b);
      const int h = static_cast<int>(in_b / static_cast<uint32_t>(p.n_qp));
      const int q0 = static_cast<int>(in_b % static_cast<uint32_t>(p.n_qp)) * 2 * A3_BQ + wg * A3_BQ;
      [[maybe_unused]] const uint32_t dbase =
          drop_key_bh(p.drop_key, static_cast<uint32_t>(b * p.H + h)) + static_cast<uint32_t>(q0 + row) * DROP_C_ROW;
      if (q0 >= p.Tq) {  // this warpgroup's tile does not exist for this item: only keep the exp turn-taking alive
        for (int j = 0; j < nk; ++j) {
          asm volatile("bar.sync %0, 256;" ::"r"(3 + wg) : "memory");
          asm volatile("bar.arrive %0, 256;" ::"r"(4 - wg) : "memory");
        }
        continue;
      }
      if (cur_b < 0 || (b != cur_b && p.key_pad != nullptr)) {   // without a mask the caps only depend on Tk
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        for (int j = wg_tid; j < n_kv; j += 128) flags[j] = 0;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        for (int kk = wg_tid; kk < n_kv * A3_BKV; kk += 128) {
          bool pad = kk >= p.Tk;
          if (!pad && p.key_pad != nullptr) pad = p.key_pad[static_cast<int64_t>(b) * p.Tk + kk] != 0;
          caps[kk] = pad ? -INFINITY : INFINITY;
          if (pad) flags[kk / A3_BKV] = 1;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        cur_b = b;
      }
      float m_run = -INFINITY;  // running reference maximum, in log2 units (score * scale * log2 e)
      float l_run = 0.0f;
      uint32_t va[32], vb[32];
      // the item's first scores
      {
        const uint32_t sbuf0 = g & 1u;
        mbar_wait(b_sfull + (wg * 2 + sbuf0) * 8, (sbuf0 ? scnt1 : scnt0) & 1u);
        if (sbuf0) ++scnt1; else ++scnt0;
        tc_fence_after_sync();
        tmem_ld32(t_tile + sbuf0 * A3_BKV, va);
        if (p.Tk > 32) tmem_ld32(t_tile + sbuf0 * A3_BKV + 32, vb);
      }
      for (int j = 0; j < nk; ++j) {
        const uint32_t sbuf = (g + static_cast<uint32_t>(j)) & 1u;
        const uint32_t t_s = t_tile + sbuf * A3_BKV;
        const int rem = p.Tk - j * A3_BKV;
        const bool two = rem > 32;                   // second 32-key chunk holds a valid key
        const bool masked = flags[j] != 0;           // warp-uniform
        const float* cap_j = caps + j * A3_BKV;
        tmem_ld_wait();
        if (masked) {
          apply_caps(va, cap_j);
          if (two) apply_caps(vb, cap_j + 32);
        }
        // ---- row maximum of this step.  Only the item's FIRST step needs it before the exponentials (it sets the
        // reference maximum).  Later steps keep the reference unless the maximum grew by more than 2^TAU (lazy
        // rescale), which is rare: they compute p = 2^(s - m_run) with the reference they have (SPECULATIVELY) while
        // the maximum is formed on the ALU pipe next to the exponentials on the XU pipe, and check afterwards; the
        // few steps that do need a rescale repair O / l and redo their exponentials before P is handed over.
        auto row_max = [&]() {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          max4(m4, va);
          if (two) max4(m4, vb);
          return fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;  // sc > 0
        };
        if (j == 0) m_run = row_max();     // (a real branch, j is warp-uniform: later steps must not wait for their maximum)
        float neg_m = (m_run == -INFINITY) ? 0.0f : -m_run;

        // ---- p = 2^(s*scale - m), row sum, bf16 P into the first 32 columns of this S buffer; the two warpgroups
        // take turns (named barriers 3 / 4) so that their exponentials do not collide on the quarter-rate XU pipe
#ifndef HRIEMO_ATTN_NO_PINGPONG
        asm volatile("bar.sync %0, 256;" ::"r"(3 + wg) : "memory");
#endif
        float l4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        uint32_t pk[16];
        exp_pack(va, sc, neg_m, l4, pk);
        if constexpr (DROP) drop_pk(pk, dbase, static_cast<uint32_t>(j) * 16u, p.drop_p8);
        tmem_st16(t_s, pk);
        if (two) {
          exp_pack(vb, sc, neg_m, l4, pk);
          if constexpr (DROP) drop_pk(pk, dbase, static_cast<uint32_t>(j) * 16u + 8u, p.drop_p8);
          tmem_st16(t_s + 16, pk);
        }
#ifndef HRIEMO_ATTN_NO_PINGPONG
        asm volatile("bar.arrive %0, 256;" ::"r"(4 - wg) : "memory");   // the other warpgroup's turn
#endif
        if (j > 0) {
          const float tile_max = row_max();
          const bool need = tile_max > m_run + A3_LAZY_TAU;
          if (__any_sync(0xffffffffu, need)) {
            // every PV of this tile must have retired: the latest one is PV(pv_issued - 1); the one before retired
            // before this step's S landed (see above), so this wait can not be answered by a stale phase
            const uint32_t kk = pv_issued - 1u;
            mbar_wait(b_pvdone + (wg * 2 + (kk & 1u)) * 8, (kk >> 1) & 1u);
            tc_fence_after_sync();
            const float alpha = need ? ex2_approx(m_run - tile_max) : 1.0f;
            if (need) m_run = tile_max;
            l_run *= alpha;
#pragma unroll 1
            for (int c = 0; c < DH / 32; ++c) {
              uint32_t v[32];
              tmem_ld32(t_o + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st32(t_o + c * 32, v);
            }
            // this step's exponentials again, against the new reference (the scores are still in registers)
            neg_m = (m_run == -INFINITY) ? 0.0f : -m_run;
            l4[0] = l4[1] = l4[2] = l4[3] = 0.0f;
            exp_pack(va, sc, neg_m, l4, pk);
            if constexpr (DROP) drop_pk(pk, dbase, static_cast<uint32_t>(j) * 16u, p.drop_p8);
            tmem_st16(t_s, pk);
            if (two) {
              exp_pack(vb, sc, neg_m, l4, pk);
              if constexpr (DROP) drop_pk(pk, dbase, static_cast<uint32_t>(j) * 16u + 8u, p.drop_p8);
              tmem_st16(t_s + 16, pk);
            }
          }
        }
        l_run += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        // ---- the next step's scores: wait for S(j+1) and issue its loads while the P stores drain
        if (j + 1 < nk) {
          const uint32_t nbuf = sbuf ^ 1u;
          mbar_wait(b_sfull + (wg * 2 + nbuf) * 8, (nbuf ? scnt1 : scnt0) & 1u);
          if (nbuf) ++scnt1; else ++scnt0;
          tc_fence_after_sync();
          tmem_ld32(t_tile + nbuf * A3_BKV, va);
          if (p.Tk - (j + 1) * A3_BKV > 32) tmem_ld32(t_tile + nbuf * A3_BKV + 32, vb);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_pfull + (wg * 2 + (pv_issued & 1u)) * 8);
        ++pv_issued;
      }
      // ---- the item's row sums go to the epilogue warpgroup; this one carries straight on with the next item
      mbar_wait(b_epiempty + wg * 8, (items_done & 1u) ^ 1u);
      xl[wg * 128 + row] = l_run;
      xm[wg * 128 + row] = m_run;
      __syncwarp();
      if (lane == 0) mbar_arrive(b_epifull + wg * 8);
      ++items_done;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <int DH>
static int launch_attention3(const hriemo_attn_args& a, cudaStream_t stream) {
  using L = Attn3Smem<DH>;
  const int d = a.H * DH;
  const int n_kv = (a.Tk + A3_BKV - 1) / A3_BKV;
  const int smem = L::dyn_bytes(n_kv);
  if (smem > 227 * 1024)
    return set_error(HRIEMO_ERR_INVALID, "attention: Tk=%d too long for the shared-memory key caps (dh=%d)",
                     a.Tk, DH);
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_bf16_2d(&tq, a.q, (uint64_t)d, (uint64_t)a.B * a.Tq, (uint64_t)a.ldq, 64, A3_BQ);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tk, a.k, (uint64_t)d, (uint64_t)a.B * a.Tk, (uint64_t)a.ldk, 64, A3_BKV);
  if (rc) return rc;
  {
    // 3-D (column, key, utterance): keys past Tk in an utterance's last tile read as ZERO instead of
    // the next utterance's rows -- P is 0 there, but 0 x NaN from a neighbouring (e.g. fully padded)
    // utterance would poison this one.
    const uint64_t dims[3] = {(uint64_t)d, (uint64_t)a.Tk, (uint64_t)a.B};
    const uint64_t pitch[2] = {(uint64_t)a.ldv, (uint64_t)a.ldv * a.Tk};
    const uint32_t box[3] = {32u, (uint32_t)A3_BKV, 1u};
    rc = make_tmap_bf16_3d_plain(&tv, a.v, dims, pitch, box, 64);
    if (rc) return rc;
  }
  CUtensorMap to;
  {
    const uint64_t dims[3] = {(uint64_t)d, (uint64_t)a.Tq, (uint64_t)a.B};
    const uint64_t pitch[2] = {(uint64_t)a.ldo, (uint64_t)a.ldo * a.Tq};
    const uint32_t box[3] = {(uint32_t)DH, 32u, 1u};
    rc = make_tmap_bf16_3d_plain(&to, a.out, dims, pitch, box);
    if (rc) return rc;
  }
  Attn3Params p;
  p.key_pad = a.key_pad;
  p.kv_steps = a.kv_steps;
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.lse = a.lse;
  p.ldo = a.ldo;
  p.B = a.B; p.H = a.H; p.Tq = a.Tq; p.Tk = a.Tk;
  p.n_kv = n_kv;
  p.n_qp = (a.Tq + 2 * A3_BQ - 1) / (2 * A3_BQ);
  p.paired = (a.Tq <= A3_BQ && a.H % 2 == 0 && !a.no_head_pairs) ? 1 : 0;
  p.items_per_b = p.paired ? a.H / 2 : a.H * p.n_qp;
  p.n_items = static_cast<int64_t>(a.B) * p.items_per_b;
  p.contiguous = n_kv <= 6 ? 1 : 0;
  if (p.n_items * n_kv * (p.paired ? 2 : 1) >= (1ll << 31))
    return set_error(HRIEMO_ERR_INVALID, "attention: too many (item, step) pairs for 32-bit step counters");
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.drop_p8 = a.drop_p8; p.drop_key = a.drop_key; p.drop_scale = a.drop_p8 ? a.drop_scale : 1.0f;
  static uint64_t attr_done = 0;
  if (device_needs_attr(&attr_done)) {
    cudaError_t e = cudaSuccess;
    const void* forms[8] = {reinterpret_cast<const void*>(attention_fwd3_kernel<DH, false, false, false>),
                            reinterpret_cast<const void*>(attention_fwd3_kernel<DH, false, true, false>),
                            reinterpret_cast<const void*>(attention_fwd3_kernel<DH, true, false, false>),
                            reinterpret_cast<const void*>(attention_fwd3_kernel<DH, true, true, false>),
                            reinterpret_cast<const void*>(attention_fwd3_kernel<DH, false, false, true>),
                            reinterpret_cast<const void*>(attention_fwd3_kernel<DH, false, true, true>),
                            reinterpret_cast<const void*>(attention_fwd3_kernel<DH, true, false, true>),
                            reinterpret_cast<const void*>(attention_fwd3_kernel<DH, true, true, true>)};
    for (int i = 0; i < 8 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(forms[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess)
      return set_error(HRIEMO_ERR_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const int64_t sms = device_sm_count();
  const unsigned grid = static_cast<unsigned>(p.n_items < sms ? p.n_items : sms);
  const bool defer = n_kv <= 2;   // short items: S look-ahead must not hold up a PV (see the kernel)
#define HRIEMO_A3_LAUNCH(PAIRED_, DEFER_)                                                                              \
  do {                                                                                                                 \
    if (p.drop_p8) attention_fwd3_kernel<DH, PAIRED_, DEFER_, true><<<grid, A3_THREADS, smem, stream>>>(tq, tk, tv, to, p); \
    else attention_fwd3_kernel<DH, PAIRED_, DEFER_, false><<<grid, A3_THREADS, smem, stream>>>(tq, tk, tv, to, p);     \
  } while (0)
  // The long-key general form has a second implementation, attention_fwd4_kernel (four column-split softmax warpgroups).
  // HRIEMO_ATTN_FWD4 = 0 / 1 selects at run time (A/B measurements in one build); default: see kFwd4Default.
  static const int fwd4_env = [] {
    const char* e = getenv("HRIEMO_ATTN_FWD4");
    return e == nullptr ? -1 : atoi(e);
  }();
  const bool use_fwd4 = !p.paired && !defer && (fwd4_env < 0 ? kFwd4Default : fwd4_env != 0) &&
                        Attn4Smem<DH>::dyn_bytes4(n_kv) <= 227 * 1024;
  // ... and a third, attention_fwd5_kernel (the O / l epilogue on its own warpgroup): HRIEMO_ATTN_FWD5 = 0 / 1.
  static const int fwd5_env = [] {
    const char* e = getenv("HRIEMO_ATTN_FWD5");
    return e == nullptr ? -1 : atoi(e);
  }();
  const bool use_fwd5 = !p.paired && !defer && !use_fwd4 && (fwd5_env < 0 ? kFwd5Default : fwd5_env != 0) &&
                        Attn5Smem<DH>::dyn_bytes5(n_kv) <= 227 * 1024;
  if (use_fwd5) {
    static uint64_t attr5_done = 0;
    if (device_needs_attr(&attr5_done)) {
      cudaError_t e = cudaFuncSetAttribute(attention_fwd5_kernel<DH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attention_fwd5_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    const int smem5 = Attn5Smem<DH>::dyn_bytes5(n_kv);
    if (p.drop_p8) attention_fwd5_kernel<DH, true><<<grid, A5_THREADS, smem5, stream>>>(tq, tk, tv, to, p);
    else attention_fwd5_kernel<DH, false><<<grid, A5_THREADS, smem5, stream>>>(tq, tk, tv, to, p);
  } else if (use_fwd4) {
    static uint64_t attr4_done = 0;
    if (device_needs_attr(&attr4_done)) {
      cudaError_t e = cudaFuncSetAttribute(attention_fwd4_kernel<DH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attention_fwd4_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    const int smem4 = Attn4Smem<DH>::dyn_bytes4(n_kv);
    if (p.drop_p8) attention_fwd4_kernel<DH, true><<<grid, A4_THREADS, smem4, stream>>>(tq, tk, tv, to, p);
    else attention_fwd4_kernel<DH, false><<<grid, A4_THREADS, smem4, stream>>>(tq, tk, tv, to, p);
  } else if (p.paired && defer) HRIEMO_A3_LAUNCH(true, true);
  else if (p.paired) HRIEMO_A3_LAUNCH(true, false);
  else if (defer) HRIEMO_A3_LAUNCH(false, true);
  else HRIEMO_A3_LAUNCH(false, false);
#undef HRIEMO_A3_LAUNCH
  return check_launch("attention_bf16");
}

}  // namespace hriemo

namespace hriemo {
// steps[b] = number of 64-key tiles up to and including the one holding the last valid key (>= 1, so a
// fully padded utterance still runs one all-masked tile and comes out NaN like torch.softmax)
__global__ void attention_kv_steps_kernel(const uint8_t* __restrict__ key_pad, int B, int Tk, int32_t* __restrict__ steps) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  int last = -1;
  for (int k = lane; k < Tk; k += 32)
    if (key_pad[static_cast<int64_t>(b) * Tk + k] == 0) last = k;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  if (lane == 0) steps[b] = last < 0 ? 1 : last / A3_BKV + 1;
}
}  // namespace hriemo

extern "C" int hriemo_attention_kv_steps(const uint8_t* key_pad, int32_t B, int32_t Tk, int32_t* steps, void* stream) {
  using namespace hriemo;
  HRIEMO_REQUIRE(key_pad && steps && B > 0 && Tk > 0, "attention_kv_steps: bad argument");
  attention_kv_steps_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(key_pad, B, Tk, steps);
  return check_launch("attention_kv_steps");
}

#ifdef HRIEMO_ATTN_TRACE
extern "C" int hriemo_debug_set_attn_trace(long long* buf) {  // trace builds only; not part of the ABI
  cudaError_t e = cudaMemcpyToSymbol(hriemo::g_attn_trace, &buf, sizeof(buf));
  return e == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int hriemo_attention_bf16(const hriemo_attn_args* a, void* stream) {
  using namespace hriemo;
  HRIEMO_REQUIRE(a != nullptr && a->q && a->k && a->v && a->out, "attention: null operand");
  HRIEMO_REQUIRE(a->B > 0 && a->H > 0 && a->Tq > 0 && a->Tk > 0, "attention: bad shape");
  HRIEMO_REQUIRE(a->B <= 65535 && a->H <= 65535, "attention: B and H must fit a grid dimension");
  HRIEMO_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0,
                 "attention: leading dimensions must be multiples of 8");
  HRIEMO_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15u) == 0, "attention: out misaligned");
  HRIEMO_REQUIRE(a->scale > 0.0f, "attention: scale must be positive");
  HRIEMO_REQUIRE(a->kv_steps == nullptr || a->key_pad != nullptr, "attention: kv_steps comes with key_pad");
  HRIEMO_REQUIRE(a->drop_p8 <= 255u && (a->drop_p8 == 0u || a->drop_scale > 0.0f), "attention: drop_p8 = round(256 p) <= 255 with drop_scale = 1 / (1 - p)");
  HRIEMO_REQUIRE(a->dh == 32 || a->dh == 64 || a->dh == 96 || a->dh == 128,
                 "attention: head dim %d not in {32,64,96,128}", a->dh);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a->dh) {
    case 32: return launch_attention3<32>(*a, s);
    case 64: return launch_attention3<64>(*a, s);
    case 96: return launch_attention3<96>(*a, s);
    default: return launch_attention3<128>(*a, s);
  }
}
