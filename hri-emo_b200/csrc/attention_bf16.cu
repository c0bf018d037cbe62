// Flash-style multi-head attention forward for sm_100a (tcgen05 + TMEM + TMA), v2:
//   O[b, tq, h*dh:(h+1)*dh] = softmax_k(Q K^T * scale + key_padding) V
//
// Persistent kernel, one CTA per SM, looping over work items (utterance, head, pair of
// 128-row query tiles).  320 threads:
//   warp 0       TMA producer: Q tiles (one per query tile), K / V^T tiles of 128 keys in a
//                2-stage ring; runs ahead across work items, so the next item's operands land
//                while the current one is still being reduced.
//   warp 1       tcgen05.mma issuer.  S_t = Q_t K^T (both operands in smem), O_t += P_t V with
//                P_t read from tensor memory (it aliases the S_t columns).  The two query tiles
//                ping-pong: while warpgroup A does softmax on S_A(j) the tensor core runs
//                PV_B / S_B(j+1), and vice versa; each K / V^T tile is loaded once for 256 rows.
//   warps 2..5   softmax warpgroup of query tile 0 \ one query row per thread (TMEM lane = row):
//   warps 6..9   softmax warpgroup of query tile 1 / raw row max, lazy running max (rescale only
//                when it grows by more than 2^8), p = ex2(fma(s, scale*log2e, -m)), P stored as
//                bf16 into TMEM, row sums in fp32; final O / l written as bf16.
//
// Replaces the scaled_dot_product_attention inside nn.MultiheadAttention at
// models/cross_modal_block_tacfn.py:74-80,85-91,98-104,111-117 and
// models/cross_modal_block.py:56-59,64-67 of the reference.
#include <math.h>
#include <stdlib.h>

#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

int attention_v1_dispatch(const hriemo_attn_args* a, cudaStream_t s);  // attention_v1.cu (A/B only)

constexpr int A2_BQ = 128;    // query rows per tile (UMMA M)
constexpr int A2_BKV = 128;   // keys per tile
constexpr int A2_THREADS = 320;
constexpr float A2_LAZY_TAU = 8.0f;  // log2 units

template <int DH>
struct Attn2Smem {
  static constexpr int QCH = (DH + 63) / 64;   // 64-column (128-byte) chunks of Q / K rows
  static constexpr int CHUNK = 128 * 128;      // [128 rows][128 B], 128B-swizzled
  static constexpr int Q_TILE = QCH * CHUNK;
  static constexpr int K_STAGE = QCH * CHUNK;
  static constexpr int V_CHUNK = DH * 128;     // [DH rows][64 keys]
  static constexpr int V_STAGE = 2 * V_CHUNK;
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = Q_OFF + 2 * Q_TILE;
  static constexpr int V_OFF = K_OFF + 2 * K_STAGE;
  static constexpr int BAR_OFF = V_OFF + 2 * V_STAGE;
  static constexpr int NUM_BARS = 18;
  static constexpr int TMEM_SLOT_OFF = BAR_OFF + NUM_BARS * 8;
  static constexpr int DYN_OFF = TMEM_SLOT_OFF + 16;   // caps[2][n_kv*128] f32, flags[2][n_kv] i32
  static int dyn_bytes(int n_kv) { return DYN_OFF + 2 * n_kv * A2_BKV * 4 + 2 * n_kv * 4 + 1024; }
};

struct Attn2Params {
  const uint8_t* key_pad;
  __nv_bfloat16* out;
  int64_t ldo;
  int B, H, Tq, Tk;
  int n_kv, n_qp;
  int64_t n_items;
  float scale_log2;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DH>
__global__ void __launch_bounds__(A2_THREADS, 1)
attention_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const Attn2Params p) {
  using L = Attn2Smem<DH>;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t TILE_COLS = 256;   // per query tile: S / P at +0 (128 cols), O at +128 (DH cols)
  constexpr uint32_t O_COL = 128;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t sQ = base + L::Q_OFF, sK = base + L::K_OFF, sV = base + L::V_OFF;
  const uint32_t bars = base + L::BAR_OFF;
  const uint32_t b_qfull = bars + 0 * 8;     // [2] per query tile
  const uint32_t b_qempty = bars + 2 * 8;    // [2]
  const uint32_t b_kfull = bars + 4 * 8;     // [2] per stage
  const uint32_t b_kempty = bars + 6 * 8;    // [2]
  const uint32_t b_vfull = bars + 8 * 8;     // [2]
  const uint32_t b_vempty = bars + 10 * 8;   // [2]
  const uint32_t b_sfull = bars + 12 * 8;    // [2] per query tile
  const uint32_t b_pfull = bars + 14 * 8;    // [2]
  const uint32_t b_pvdone = bars + 16 * 8;   // [2]
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::TMEM_SLOT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = p.n_kv;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_qfull + s * 8, 1);
      mbar_init(b_qempty + s * 8, 1);
      mbar_init(b_kfull + s * 8, 1);
      mbar_init(b_kempty + s * 8, 1);
      mbar_init(b_vfull + s * 8, 1);
      mbar_init(b_vempty + s * 8, 1);
      mbar_init(b_sfull + s * 8, 1);
      mbar_init(b_pfull + s * 8, 128);
      mbar_init(b_pvdone + s * 8, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(base + L::TMEM_SLOT_OFF);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t qcnt[2] = {0, 0};
      uint32_t g = 0;  // running K/V tile index -> ring stage and phase
      for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int qp = static_cast<int>(item % p.n_qp);
        const int64_t bh = item / p.n_qp;
        const int h = static_cast<int>(bh % p.H);
        const int b = static_cast<int>(bh / p.H);
        const int q0 = qp * 2 * A2_BQ;
        for (int t = 0; t < 2; ++t) {
          if (q0 + t * A2_BQ >= p.Tq) break;
          mbar_wait(b_qempty + t * 8, (qcnt[t] & 1) ^ 1);
          mbar_arrive_expect_tx(b_qfull + t * 8, L::Q_TILE);
          for (int c = 0; c < L::QCH; ++c)
            tma_load_2d(&tm_q, b_qfull + t * 8, sQ + t * L::Q_TILE + c * L::CHUNK, h * DH + c * 64,
                        b * p.Tq + q0 + t * A2_BQ);
          ++qcnt[t];
        }
        for (int j = 0; j < n_kv; ++j, ++g) {
          const uint32_t s = g & 1u, par = (g >> 1) & 1u;
          mbar_wait(b_kempty + s * 8, par ^ 1);
          mbar_arrive_expect_tx(b_kfull + s * 8, L::K_STAGE);
          for (int c = 0; c < L::QCH; ++c)
            tma_load_2d(&tm_k, b_kfull + s * 8, sK + s * L::K_STAGE + c * L::CHUNK, h * DH + c * 64,
                        b * p.Tk + j * A2_BKV);
          mbar_wait(b_vempty + s * 8, par ^ 1);
          mbar_arrive_expect_tx(b_vfull + s * 8, L::V_STAGE);
          for (int c = 0; c < 2; ++c)
            tma_load_2d(&tm_v, b_vfull + s * 8, sV + s * L::V_STAGE + c * L::V_CHUNK, j * A2_BKV + c * 64,
                        (b * p.H + h) * DH);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(A2_BQ, A2_BKV);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(A2_BQ, DH);
      uint32_t qcnt[2] = {0, 0}, pcnt[2] = {0, 0};
      uint32_t g = 0;
      auto issue_s = [&](int t, uint32_t stage) {
#pragma unroll
        for (int st = 0; st < DH / 16; ++st) {
          const uint32_t off = (st >> 2) * L::CHUNK + (st & 3) * 32;
          umma_bf16(tmem_base + t * TILE_COLS, umma_desc_sw128(sQ + t * L::Q_TILE + off),
                    umma_desc_sw128(sK + stage * L::K_STAGE + off), idesc_s, st != 0);
        }
        umma_commit(b_sfull + t * 8);
      };
      for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int qp = static_cast<int>(item % p.n_qp);
        const int nt = (qp * 2 * A2_BQ + A2_BQ < p.Tq) ? 2 : 1;  // active query tiles
        // ---- S_t(0)
        {
          const uint32_t s = g & 1u, par = (g >> 1) & 1u;
          for (int t = 0; t < nt; ++t) mbar_wait(b_qfull + t * 8, qcnt[t] & 1);
          mbar_wait(b_kfull + s * 8, par);
          tc_fence_after_sync();
          for (int t = 0; t < nt; ++t) {
            issue_s(t, s);
            if (n_kv == 1) umma_commit(b_qempty + t * 8);
          }
          umma_commit(b_kempty + s * 8);
        }
        for (int j = 0; j < n_kv; ++j, ++g) {
          const uint32_t s = g & 1u, par = (g >> 1) & 1u;
          const bool more = j + 1 < n_kv;
          const uint32_t s2 = (g + 1) & 1u, par2 = ((g + 1) >> 1) & 1u;
          const int rem = p.Tk - j * A2_BKV;  // keys left from this tile on (> 0)
          mbar_wait(b_vfull + s * 8, par);
          if (more) mbar_wait(b_kfull + s2 * 8, par2);
          for (int t = 0; t < nt; ++t) {
            mbar_wait(b_pfull + t * 8, pcnt[t] & 1);
            ++pcnt[t];
            tc_fence_after_sync();
#pragma unroll
            for (int st = 0; st < A2_BKV / 16; ++st) {
              if (st * 16 < rem) {  // P is zero and V^T zero-filled beyond Tk: skip those K-steps
                umma_bf16_ts(tmem_base + t * TILE_COLS + O_COL, tmem_base + t * TILE_COLS + st * 8,
                             umma_desc_sw128(sV + s * L::V_STAGE + (st >> 2) * L::V_CHUNK + (st & 3) * 32),
                             idesc_pv, (j | st) != 0);
              }
            }
            umma_commit(b_pvdone + t * 8);
            if (more) {
              issue_s(t, s2);  // overwrites S_t / P_t: ordered after PV_t(j) by in-order MMA execution
              if (j + 2 == n_kv) umma_commit(b_qempty + t * 8);
            }
          }
          umma_commit(b_vempty + s * 8);
          if (more) umma_commit(b_kempty + s2 * 8);
        }
        for (int t = 0; t < nt; ++t) ++qcnt[t];
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int wg = (warp - 2) >> 2;          // query tile handled by this warpgroup
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;          // query row inside the tile == TMEM lane
    const int wg_tid = threadIdx.x - 64 - wg * 128;
    float* caps = reinterpret_cast<float*>(base_ptr + L::DYN_OFF) + wg * n_kv * A2_BKV;
    int* flags = reinterpret_cast<int*>(base_ptr + L::DYN_OFF + 2 * n_kv * A2_BKV * 4) + wg * n_kv;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_sel + wg * TILE_COLS;
    const uint32_t t_o = t_s + O_COL;
    uint32_t scnt = 0, pvcnt = 0;
    int cur_b = -1;

    for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int qp = static_cast<int>(item % p.n_qp);
      const int64_t bh = item / p.n_qp;
      const int h = static_cast<int>(bh % p.H);
      const int b = static_cast<int>(bh / p.H);
      const int q0 = qp * 2 * A2_BQ + wg * A2_BQ;
      if (q0 >= p.Tq) continue;  // this warpgroup's tile does not exist for this item

      if (b != cur_b) {
        // key caps: +inf for valid keys, -inf for PAD keys and keys >= Tk.  score = fminf(s, cap):
        // fminf returns the non-NaN operand, so whatever a masked column holds (rows of the next
        // utterance inside the 128-key box, possibly NaN) becomes exactly -inf.
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        for (int j = wg_tid; j < n_kv; j += 128) flags[j] = 0;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        for (int kk = wg_tid; kk < n_kv * A2_BKV; kk += 128) {
          bool pad = kk >= p.Tk;
          if (!pad && p.key_pad != nullptr) pad = p.key_pad[static_cast<int64_t>(b) * p.Tk + kk] != 0;
          caps[kk] = pad ? -INFINITY : INFINITY;
          if (pad) flags[kk / A2_BKV] = 1;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        cur_b = b;
      }

      float m_run = -INFINITY;  // running reference maximum, in log2 units (score * scale * log2 e)
      float l_run = 0.0f;
      for (int j = 0; j < n_kv; ++j) {
        const int rem = p.Tk - j * A2_BKV;
        const int nch = rem >= A2_BKV ? 4 : (rem + 31) / 32;  // 32-key chunks holding any valid key
        const bool masked = flags[j] != 0;                    // warp-uniform
        const float* cap_j = caps + j * A2_BKV;
        mbar_wait(b_sfull + wg * 8, scnt & 1);
        ++scnt;
        tc_fence_after_sync();

        // ---- pass 1: raw row maximum
        float mx = -INFINITY;
        for (int c = 0; c < nch; ++c) {
          uint32_t v[32];
          tmem_ld32(t_s + c * 32, v);
          tmem_ld_wait();
          if (masked) {
            const float4* cp = reinterpret_cast<const float4*>(cap_j + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 ck = cp[i];
              mx = fmaxf(mx, fminf(__uint_as_float(v[i * 4 + 0]), ck.x));
              mx = fmaxf(mx, fminf(__uint_as_float(v[i * 4 + 1]), ck.y));
              mx = fmaxf(mx, fminf(__uint_as_float(v[i * 4 + 2]), ck.z));
              mx = fmaxf(mx, fminf(__uint_as_float(v[i * 4 + 3]), ck.w));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        const float tile_max = mx * p.scale_log2;  // scale > 0; -inf stays -inf

        // ---- lazy running maximum; O / l rescale only when the maximum grew by more than 2^TAU
        if (j == 0) {
          m_run = tile_max;
        } else {
          const bool need = tile_max > m_run + A2_LAZY_TAU;
          mbar_wait(b_pvdone + wg * 8, pvcnt & 1);  // PV(j-1) retired: O is stable
          ++pvcnt;
          tc_fence_after_sync();
          if (__any_sync(0xffffffffu, need)) {
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

        // ---- pass 2: p = 2^(s*scale - m), row sum, bf16 P into TMEM (aliases the S columns already read)
        float l_add = 0.0f;
        for (int c = 0; c < nch; ++c) {
          uint32_t v[32];
          tmem_ld32(t_s + c * 32, v);
          tmem_ld_wait();
          float e[32];
          if (masked) {
            const float4* cp = reinterpret_cast<const float4*>(cap_j + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 ck = cp[i];
              e[i * 4 + 0] = ex2_approx(fmaf(fminf(__uint_as_float(v[i * 4 + 0]), ck.x), p.scale_log2, neg_m));
              e[i * 4 + 1] = ex2_approx(fmaf(fminf(__uint_as_float(v[i * 4 + 1]), ck.y), p.scale_log2, neg_m));
              e[i * 4 + 2] = ex2_approx(fmaf(fminf(__uint_as_float(v[i * 4 + 2]), ck.z), p.scale_log2, neg_m));
              e[i * 4 + 3] = ex2_approx(fmaf(fminf(__uint_as_float(v[i * 4 + 3]), ck.w), p.scale_log2, neg_m));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, neg_m));
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            l_add += e[2 * i] + e[2 * i + 1];
            pk[i] = pack_bf16(e[2 * i], e[2 * i + 1]);
          }
          tmem_st16(t_s + c * 16, pk);
        }
        l_run += l_add;
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive(b_pfull + wg * 8);
      }

      // ---- epilogue: O / l -> bf16 rows of the [B*Tq, H*dh] output
      mbar_wait(b_pvdone + wg * 8, pvcnt & 1);
      ++pvcnt;
      tc_fence_after_sync();
      const float inv_l = 1.0f / l_run;  // l == 0 (every key masked) -> inf -> NaN like torch.softmax
      const bool row_ok = q0 + r < p.Tq;
      __nv_bfloat16* orow = p.out + (static_cast<int64_t>(b) * p.Tq + q0 + r) * p.ldo + h * DH;
#pragma unroll 1
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t_o + c * 32, v);
        tmem_ld_wait();
        if (row_ok) {
          uint4* d4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(v[g4 * 8 + 0]) * inv_l, __uint_as_float(v[g4 * 8 + 1]) * inv_l);
            w.y = pack_bf16(__uint_as_float(v[g4 * 8 + 2]) * inv_l, __uint_as_float(v[g4 * 8 + 3]) * inv_l);
            w.z = pack_bf16(__uint_as_float(v[g4 * 8 + 4]) * inv_l, __uint_as_float(v[g4 * 8 + 5]) * inv_l);
            w.w = pack_bf16(__uint_as_float(v[g4 * 8 + 6]) * inv_l, __uint_as_float(v[g4 * 8 + 7]) * inv_l);
            d4[g4] = w;
          }
        }
        __syncwarp();
      }
      tc_fence_before_sync();  // O reads are complete before the next item's p_full arrival lets PV overwrite O
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
static int launch_attention2(const hriemo_attn_args& a, cudaStream_t stream) {
  using L = Attn2Smem<DH>;
  const int d = a.H * DH;
  const int n_kv = (a.Tk + A2_BKV - 1) / A2_BKV;
  const int smem = L::dyn_bytes(n_kv);
  if (smem > 227 * 1024)
    return set_error(HRIEMO_ERR_INVALID, "attention: Tk=%d too long for the shared-memory key caps (dh=%d)",
                     a.Tk, DH);
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_bf16_2d(&tq, a.q, (uint64_t)d, (uint64_t)a.B * a.Tq, (uint64_t)a.ldq, 64, A2_BQ);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tk, a.k, (uint64_t)d, (uint64_t)a.B * a.Tk, (uint64_t)a.ldk, 64, A2_BKV);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tv, a.vt, (uint64_t)a.Tk, (uint64_t)a.B * d, (uint64_t)a.Tk_pad, 64, DH);
  if (rc) return rc;
  Attn2Params p;
  p.key_pad = a.key_pad;
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.ldo = a.ldo;
  p.B = a.B; p.H = a.H; p.Tq = a.Tq; p.Tk = a.Tk;
  p.n_kv = n_kv;
  p.n_qp = (a.Tq + 2 * A2_BQ - 1) / (2 * A2_BQ);
  p.n_items = static_cast<int64_t>(a.B) * a.H * p.n_qp;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd2_kernel<DH>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess)
      return set_error(HRIEMO_ERR_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t sms = device_sm_count();
  const unsigned grid = static_cast<unsigned>(p.n_items < sms ? p.n_items : sms);
  attention_fwd2_kernel<DH><<<grid, A2_THREADS, smem, stream>>>(tq, tk, tv, p);
  return check_launch("attention_bf16");
}

}  // namespace hriemo

extern "C" int hriemo_attention_bf16(const hriemo_attn_args* a, void* stream) {
  using namespace hriemo;
  HRIEMO_REQUIRE(a != nullptr && a->q && a->k && a->vt && a->out, "attention: null operand");
  HRIEMO_REQUIRE(a->B > 0 && a->H > 0 && a->Tq > 0 && a->Tk > 0, "attention: bad shape");
  HRIEMO_REQUIRE(a->B <= 65535 && a->H <= 65535, "attention: B and H must fit a grid dimension");
  HRIEMO_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldo % 8 == 0 && a->Tk_pad % 8 == 0 &&
                     a->Tk_pad >= a->Tk,
                 "attention: leading dimensions must be multiples of 8");
  HRIEMO_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15u) == 0, "attention: out misaligned");
  HRIEMO_REQUIRE(a->scale > 0.0f, "attention: scale must be positive");
  HRIEMO_REQUIRE(a->dh == 32 || a->dh == 64 || a->dh == 96 || a->dh == 128,
                 "attention: head dim %d not in {32,64,96,128}", a->dh);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const bool use_v1 = getenv("HRIEMO_ATTN_V1") != nullptr;  // A/B switch while v2 is validated
  if (use_v1) return attention_v1_dispatch(a, s);
  switch (a->dh) {
    case 32: return launch_attention2<32>(*a, s);
    case 64: return launch_attention2<64>(*a, s);
    case 96: return launch_attention2<96>(*a, s);
    default: return launch_attention2<128>(*a, s);
  }
}
