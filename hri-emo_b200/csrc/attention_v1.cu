// Flash-style multi-head attention forward for sm_100a (tcgen05 + TMEM + TMA):
//   O[b, tq, h*dh:(h+1)*dh] = softmax_k(Q K^T * scale + key_padding) V
// One CTA per (q-tile of 128 rows, head, utterance).
//   warp 0      : TMA producer (Q once; K / V^T tiles of 128 keys, double-buffered)
//   warp 1      : TMEM allocator + tcgen05.mma issuer (S = Q K^T, O += P V)
//   warps 2..5  : softmax + epilogue, one query row per thread (TMEM lane = row),
//                 online softmax in registers, P written to shared memory as the
//                 128B-swizzled K-major A operand of the PV MMA.
// S is double-buffered in TMEM so S(j+1) is computed while softmax(j) runs.
// The running maximum is updated lazily (only when it grows by more than 2^8),
// which keeps the O rescale off the common path and is exact in real arithmetic.
//
// Replaces the scaled_dot_product_attention inside nn.MultiheadAttention at
// models/cross_modal_block_tacfn.py:74-80,85-91,98-104,111-117 and
// models/cross_modal_block.py:56-59,64-67 of the reference.
#include <math.h>

#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

constexpr int ATT_BQ = 128;   // query rows per CTA (UMMA M)
constexpr int ATT_BKV = 128;  // keys per tile (UMMA N of S, K of PV)
constexpr int ATT_THREADS = 192;
constexpr float ATT_LAZY_TAU = 8.0f;  // log2 units

template <int DH>
struct AttnSmem {
  static constexpr int QCH = (DH + 63) / 64;               // 64-column chunks of Q / K
  static constexpr int CHUNK = 128 * 128;                  // [128 rows][128 B]
  static constexpr int Q_BYTES = QCH * CHUNK;
  static constexpr int K_STAGE = QCH * CHUNK;
  static constexpr int V_CHUNK = DH * 128;                 // [DH rows][64 keys]
  static constexpr int V_STAGE = 2 * V_CHUNK;
  static constexpr int V_STAGES = (DH > 96) ? 1 : 2;
  static constexpr int P_BUF = 2 * CHUNK;                  // [128 rows][128 keys] bf16
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = Q_OFF + Q_BYTES;
  static constexpr int V_OFF = K_OFF + 2 * K_STAGE;
  static constexpr int P_OFF = V_OFF + V_STAGES * V_STAGE;
  static constexpr int BAR_OFF = P_OFF + 2 * P_BUF;
  // q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], s_empty[2], p_full[2], pv_done
  static constexpr int NUM_BARS = 16;
  static constexpr int TMEM_SLOT_OFF = BAR_OFF + NUM_BARS * 8;
  static constexpr int MASK_OFF = TMEM_SLOT_OFF + 16;
  static int dyn_bytes(int n_kv_tiles) { return MASK_OFF + n_kv_tiles * ATT_BKV * 4 + 1024; }
};

struct AttnKernelParams {
  const uint8_t* key_pad;
  __nv_bfloat16* out;
  int64_t ldo;
  int B, H, Tq, Tk;
  int n_kv_tiles;
  float scale_log2;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DH>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const AttnKernelParams p) {
  using L = AttnSmem<DH>;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t S_COL0 = 0, O_COL = 256;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t sQ = base + L::Q_OFF, sK = base + L::K_OFF, sV = base + L::V_OFF, sP = base + L::P_OFF;
  const uint32_t bars = base + L::BAR_OFF;
  const uint32_t b_qfull = bars + 0 * 8;
  const uint32_t b_kfull = bars + 1 * 8;    // [2]
  const uint32_t b_kempty = bars + 3 * 8;   // [2]
  const uint32_t b_vfull = bars + 5 * 8;    // [2]
  const uint32_t b_vempty = bars + 7 * 8;   // [2]
  const uint32_t b_sfull = bars + 9 * 8;    // [2]
  const uint32_t b_sempty = bars + 11 * 8;  // [2]
  const uint32_t b_pfull = bars + 13 * 8;   // [2]
  const uint32_t b_pvdone = bars + 15 * 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::TMEM_SLOT_OFF);
  float* madd = reinterpret_cast<float*>(base_ptr + L::MASK_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(b_qfull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_kfull + s * 8, 1);
      mbar_init(b_kempty + s * 8, 1);
      mbar_init(b_vfull + s * 8, 1);
      mbar_init(b_vempty + s * 8, 1);
      mbar_init(b_sfull + s * 8, 1);
      mbar_init(b_sempty + s * 8, 128);
      mbar_init(b_pfull + s * 8, 128);
    }
    mbar_init(b_pvdone, 1);
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
      mbar_arrive_expect_tx(b_qfull, L::Q_BYTES);
      for (int c = 0; c < L::QCH; ++c)
        tma_load_2d(&tm_q, b_qfull, sQ + c * L::CHUNK, h * DH + c * 64, b * p.Tq + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j & 1;
        mbar_wait(b_kempty + ks * 8, ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(b_kfull + ks * 8, L::K_STAGE);
        for (int c = 0; c < L::QCH; ++c)
          tma_load_2d(&tm_k, b_kfull + ks * 8, sK + ks * L::K_STAGE + c * L::CHUNK, h * DH + c * 64,
                      b * p.Tk + j * ATT_BKV);
        const int vs = j % L::V_STAGES;
        mbar_wait(b_vempty + vs * 8, ((j / L::V_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(b_vfull + vs * 8, L::V_STAGE);
        for (int c = 0; c < 2; ++c)
          tma_load_2d(&tm_v, b_vfull + vs * 8, sV + vs * L::V_STAGE + c * L::V_CHUNK,
                      j * ATT_BKV + c * 64, (b * p.H + h) * DH);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BQ, ATT_BKV);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BQ, DH);
      auto issue_s = [&](int j) {
        const int sb = j & 1;
        mbar_wait(b_kfull + sb * 8, (j >> 1) & 1);
        mbar_wait(b_sempty + sb * 8, ((j >> 1) & 1) ^ 1);
        tc_fence_after_sync();
#pragma unroll
        for (int st = 0; st < DH / 16; ++st) {
          const uint32_t off = (st >> 2) * L::CHUNK + (st & 3) * 32;
          umma_bf16(tmem_base + S_COL0 + sb * ATT_BKV, umma_desc_sw128(sQ + off),
                    umma_desc_sw128(sK + sb * L::K_STAGE + off), idesc_s, st != 0);
        }
        umma_commit(b_kempty + sb * 8);
        umma_commit(b_sfull + sb * 8);
      };
      mbar_wait(b_qfull, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        const int pb = j & 1;
        const int vs = j % L::V_STAGES;
        mbar_wait(b_pfull + pb * 8, (j >> 1) & 1);
        mbar_wait(b_vfull + vs * 8, (j / L::V_STAGES) & 1);
        tc_fence_after_sync();
        const int rem = p.Tk - j * ATT_BKV;  // keys left in this tile (> 0)
#pragma unroll
        for (int st = 0; st < ATT_BKV / 16; ++st) {
          if (st * 16 < rem) {  // P is zero (and V^T zero-filled) beyond Tk: skip those K-steps
            const uint32_t kc = st >> 2, kk = st & 3;
            umma_bf16(tmem_base + O_COL, umma_desc_sw128(sP + pb * L::P_BUF + kc * L::CHUNK + kk * 32),
                      umma_desc_sw128(sV + vs * L::V_STAGE + kc * L::V_CHUNK + kk * 32), idesc_pv,
                      (j | st) != 0);
          }
        }
        umma_commit(b_vempty + vs * 8);
        umma_commit(b_pvdone);
      }
    }
  } else {
    // ===================== softmax + epilogue (128 threads) =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const int st_tid = threadIdx.x - 64;
    // key cap: +inf for valid keys, -inf for PAD keys and keys >= Tk.  score = fminf(s*scale, cap):
    // fminf returns the non-NaN operand, so whatever a masked column holds (rows of the next
    // utterance inside the 128-key box, possibly NaN) becomes exactly -inf.
    for (int kk = st_tid; kk < n_kv * ATT_BKV; kk += 128) {
      bool pad = kk >= p.Tk;
      if (!pad && p.key_pad != nullptr) pad = p.key_pad[static_cast<int64_t>(b) * p.Tk + kk] != 0;
      madd[kk] = pad ? -INFINITY : INFINITY;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");

    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_o = tmem_base + lane_sel + O_COL;
    float m_run = -INFINITY;
    float l_run = 0.0f;

    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      const uint32_t t_s = tmem_base + lane_sel + S_COL0 + sb * ATT_BKV;
      const int rem = p.Tk - j * ATT_BKV;
      const int nch = rem >= ATT_BKV ? 4 : (rem + 31) / 32;  // 32-key chunks with any valid key
      const float* madd_j = madd + j * ATT_BKV;
      mbar_wait(b_sfull + sb * 8, (j >> 1) & 1);
      tc_fence_after_sync();

      // ---- pass 1: row maximum of the scaled, masked scores
      float tile_max = -INFINITY;
      for (int c = 0; c < nch; ++c) {
        uint32_t v[32];
        tmem_ld32(t_s + c * 32, v);
        tmem_ld_wait();
        const float4* mp = reinterpret_cast<const float4*>(madd_j + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 mk = mp[i];
          tile_max = fmaxf(tile_max, fminf(__uint_as_float(v[i * 4 + 0]) * p.scale_log2, mk.x));
          tile_max = fmaxf(tile_max, fminf(__uint_as_float(v[i * 4 + 1]) * p.scale_log2, mk.y));
          tile_max = fmaxf(tile_max, fminf(__uint_as_float(v[i * 4 + 2]) * p.scale_log2, mk.z));
          tile_max = fmaxf(tile_max, fminf(__uint_as_float(v[i * 4 + 3]) * p.scale_log2, mk.w));
        }
      }

      // ---- running maximum (lazy) and O / l rescale
      if (j == 0) {
        m_run = tile_max;
      } else {
        const bool need = tile_max > m_run + ATT_LAZY_TAU;
        mbar_wait(b_pvdone, (j - 1) & 1);  // PV(j-1) retired: O is stable, P[sb] is free again
        tc_fence_after_sync();
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? fast_exp2(m_run - tile_max) : 1.0f;
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
          tmem_st_wait();
        }
      }
      const float m_eff = (m_run == -INFINITY) ? 0.0f : m_run;

      // ---- pass 2: P = exp2(s - m), row sum, bf16 P tile into swizzled smem
      float l_add = 0.0f;
      const uint32_t p_row = sP + sb * L::P_BUF + r * 128;
      for (int c = 0; c < nch; ++c) {
        uint32_t v[32];
        tmem_ld32(t_s + c * 32, v);
        tmem_ld_wait();
        const float4* mp = reinterpret_cast<const float4*>(madd_j + c * 32);
        float e[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 mk = mp[i];
          e[i * 4 + 0] = fast_exp2(fminf(__uint_as_float(v[i * 4 + 0]) * p.scale_log2, mk.x) - m_eff);
          e[i * 4 + 1] = fast_exp2(fminf(__uint_as_float(v[i * 4 + 1]) * p.scale_log2, mk.y) - m_eff);
          e[i * 4 + 2] = fast_exp2(fminf(__uint_as_float(v[i * 4 + 2]) * p.scale_log2, mk.z) - m_eff);
          e[i * 4 + 3] = fast_exp2(fminf(__uint_as_float(v[i * 4 + 3]) * p.scale_log2, mk.w) - m_eff);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) l_add += e[i];
        const uint32_t chunk_base = p_row + (c >> 1) * L::CHUNK;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t slot = static_cast<uint32_t>(((c & 1) * 4 + g) ^ (r & 7));
          const uint32_t w0 = pack_bf16(e[g * 8 + 0], e[g * 8 + 1]);
          const uint32_t w1 = pack_bf16(e[g * 8 + 2], e[g * 8 + 3]);
          const uint32_t w2 = pack_bf16(e[g * 8 + 4], e[g * 8 + 5]);
          const uint32_t w3 = pack_bf16(e[g * 8 + 6], e[g * 8 + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(chunk_base + slot * 16), "r"(w0),
                       "r"(w1), "r"(w2), "r"(w3)
                       : "memory");
        }
      }
      l_run += l_add;

      tc_fence_before_sync();
      mbar_arrive(b_sempty + sb * 8);  // S[sb] fully consumed
      fence_proxy_async_smem();        // P visible to the tensor-core (async) proxy
      mbar_arrive(b_pfull + sb * 8);
    }

    // ---- epilogue: O / l -> bf16
    mbar_wait(b_pvdone, (n_kv - 1) & 1);
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
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
          w.y = pack_bf16(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
          w.z = pack_bf16(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
          w.w = pack_bf16(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
          d4[g] = w;
        }
      }
      __syncwarp();
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
static int launch_attention(const hriemo_attn_args& a, cudaStream_t stream) {
  using L = AttnSmem<DH>;
  const int d = a.H * DH;
  const int n_kv = (a.Tk + ATT_BKV - 1) / ATT_BKV;
  const int smem = L::dyn_bytes(n_kv);
  if (smem > 227 * 1024)
    return set_error(HRIEMO_ERR_INVALID, "attention: Tk=%d too long for the shared-memory mask (dh=%d)",
                     a.Tk, DH);
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_bf16_2d(&tq, a.q, (uint64_t)d, (uint64_t)a.B * a.Tq, (uint64_t)a.ldq, 64, ATT_BQ);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tk, a.k, (uint64_t)d, (uint64_t)a.B * a.Tk, (uint64_t)a.ldk, 64, ATT_BKV);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tv, a.vt, (uint64_t)a.Tk, (uint64_t)a.B * d, (uint64_t)a.Tk_pad, 64, DH);
  if (rc) return rc;
  AttnKernelParams p;
  p.key_pad = a.key_pad;
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.ldo = a.ldo;
  p.B = a.B; p.H = a.H; p.Tq = a.Tq; p.Tk = a.Tk;
  p.n_kv_tiles = n_kv;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<DH>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess)
      return set_error(HRIEMO_ERR_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_bytes = 227 * 1024;
  }
  dim3 grid((a.Tq + ATT_BQ - 1) / ATT_BQ, a.H, a.B);
  attention_fwd_kernel<DH><<<grid, ATT_THREADS, smem, stream>>>(tq, tk, tv, p);
  return check_launch("attention_bf16");
}

}  // namespace hriemo

namespace hriemo {
int attention_v1_dispatch(const hriemo_attn_args* a, cudaStream_t s) {
  switch (a->dh) {
    case 32: return launch_attention<32>(*a, s);
    case 64: return launch_attention<64>(*a, s);
    case 96: return launch_attention<96>(*a, s);
    case 128: return launch_attention<128>(*a, s);
    default: return set_error(HRIEMO_ERR_INVALID, "attention: head dim %d not in {32,64,96,128}", a->dh);
  }
}
}  // namespace hriemo
