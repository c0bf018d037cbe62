// Backward of hriemo_attention_bf16 (the encoder's self- and cross-attention, models/cross_modal_block_tacfn.py:74-80,
// 85-91, 98-104, 111-117) for the training step (BASELINE config 5).
//
//   P  = exp(scale * Q K^T - LSE)            (rebuilt from the forward's log-sum-exp, no running maximum needed)
//   dV = P^T dO,  dP = dO V^T,  dS = scale * P o (dP - D),  D = rowsum(dO o O),  dQ = dS K,  dK = dS^T Q
//
// NOT the final form: the contractions run on the tensor cores through warp-level
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with operands staged in padded shared memory, NOT through
// tcgen05/TMEM/TMA like the forward kernel (attention_bf16.cu); DESIGN.md sec. 8 describes the tcgen05 form that
// replaces it.  Two deterministic passes instead of atomics: one CTA per (utterance, head, 64-key tile) walks the
// query tiles and owns dK / dV of its keys; one CTA per (utterance, head, 64-query tile) walks the key tiles and
// owns dQ of its queries.  S and dP are therefore computed twice (7 tile GEMMs per tile pair instead of 5).
//
// impl == 1 replaces every mma.sync by plain fp32 FMA loops over the same shared-memory tiles (same
// thread-to-element mapping): the slow reference form the tests compare the tensor-core form with.
#include <cuda_bf16.h>

#include "host_common.h"

namespace hriemo {

using bf = __nv_bfloat16;

constexpr int AB_T = 64;          // queries per tile and keys per tile
constexpr int AB_LDT = AB_T + 8;  // padded pitch of a [.][64] tile: 36 words, rows g = 0..7 fall on distinct banks
constexpr int AB_THREADS = 256;   // 8 warps: 4 row blocks of 16 x 2 column halves

// acc[16 x NT*8] += A[16 x K] . B[NT*8 x K]^T with A, B row-major bf16 in shared memory (A at the warp's first row,
// B at the warp's first output column).  Fragment layout of mma.m16n8k16 (g = lane / 4, t = lane % 4):
//   A regs: (g, 2t..2t+1) (g+8, 2t..) (g, 2t+8..) (g+8, 2t+8..);  B regs: (k = 2t..2t+1, n = g) (k = 2t+8.., n = g);
//   C regs: (g, 2t) (g, 2t+1) (g+8, 2t) (g+8, 2t+1).
template <int NT, bool USE_MMA>
__device__ __forceinline__ void warp_gemm(float (&acc)[NT][4], const bf* __restrict__ A, int lda,
                                          const bf* __restrict__ B, int ldb, int K, int lane) {
  const int g = lane >> 2, t = lane & 3;
  if constexpr (USE_MMA) {
    for (int k0 = 0; k0 < K; k0 += 16) {
      const uint32_t a0 = *reinterpret_cast<const uint32_t*>(A + g * lda + k0 + 2 * t);
      const uint32_t a1 = *reinterpret_cast<const uint32_t*>(A + (g + 8) * lda + k0 + 2 * t);
      const uint32_t a2 = *reinterpret_cast<const uint32_t*>(A + g * lda + k0 + 2 * t + 8);
      const uint32_t a3 = *reinterpret_cast<const uint32_t*>(A + (g + 8) * lda + k0 + 2 * t + 8);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(B + (nt * 8 + g) * ldb + k0 + 2 * t);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(B + (nt * 8 + g) * ldb + k0 + 2 * t + 8);
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
            : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      }
    }
  } else {
    for (int k = 0; k < K; ++k) {
      const float x0 = __bfloat162float(A[g * lda + k]), x1 = __bfloat162float(A[(g + 8) * lda + k]);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float y0 = __bfloat162float(B[(nt * 8 + 2 * t) * ldb + k]);
        const float y1 = __bfloat162float(B[(nt * 8 + 2 * t + 1) * ldb + k]);
        acc[nt][0] = fmaf(x0, y0, acc[nt][0]);
        acc[nt][1] = fmaf(x0, y1, acc[nt][1]);
        acc[nt][2] = fmaf(x1, y0, acc[nt][2]);
        acc[nt][3] = fmaf(x1, y1, acc[nt][3]);
      }
    }
  }
}

// Rows [row0, row0 + 64) of one head of a row-major bf16 matrix, in two steps so that the global loads of the next
// tile are in flight while the current one is being multiplied: fetch_tile (global -> registers) and stash_tile
// (registers -> shared memory, as they are ([64][DH + 8]) and, when dstT != nullptr, transposed ([DH][72])).  Rows >=
// row_end are zero.  Consecutive lanes take consecutive ROWS of one 8-column chunk: the transposed 2-byte stores of a
// warp then fall into 16 consecutive words (no bank conflict; lanes with consecutive columns would all hit one bank,
// 8 * 72 elements apart), and the 16-byte stores of the plain tile are conflict-free through the padded pitch.
template <int DH>
struct TileRegs {
  static constexpr int N = (AB_T * (DH / 8)) / AB_THREADS;   // 16-byte chunks per thread: DH / 32
  uint4 v[N];
};

template <int DH>
__device__ __forceinline__ void fetch_tile(TileRegs<DH>& regs, const bf* __restrict__ src, int64_t ld, int64_t row0,
                                           int64_t row_end) {
#pragma unroll
  for (int j = 0; j < TileRegs<DH>::N; ++j) {
    const int idx = threadIdx.x + j * AB_THREADS;
    const int r = idx % AB_T, c = (idx / AB_T) * 8;
    regs.v[j] = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < row_end) regs.v[j] = __ldg(reinterpret_cast<const uint4*>(src + (row0 + r) * ld + c));
  }
}

template <int DH>
__device__ __forceinline__ void stash_tile(const TileRegs<DH>& regs, bf* __restrict__ dst, bf* __restrict__ dstT) {
  constexpr int LDH = DH + 8;
#pragma unroll
  for (int j = 0; j < TileRegs<DH>::N; ++j) {
    const int idx = threadIdx.x + j * AB_THREADS;
    const int r = idx % AB_T, c = (idx / AB_T) * 8;
    *reinterpret_cast<uint4*>(dst + r * LDH + c) = regs.v[j];
    if (dstT != nullptr) {
      const bf* e = reinterpret_cast<const bf*>(&regs.v[j]);
#pragma unroll
      for (int k = 0; k < 8; ++k) dstT[(c + k) * AB_LDT + r] = e[k];
    }
  }
}

template <int DH>
__device__ __forceinline__ void load_tile(const bf* __restrict__ src, int64_t ld, int64_t row0, int64_t row_end,
                                          bf* __restrict__ dst, bf* __restrict__ dstT) {
  TileRegs<DH> regs;
  fetch_tile<DH>(regs, src, ld, row0, row_end);
  stash_tile<DH>(regs, dst, dstT);
}

struct AttnBwdParams {
  const bf *q, *k, *v, *d_out;
  int64_t ldq, ldk, ldv, lddo;
  const float *lse, *dsum;        // [B, H, Tq]
  const uint8_t* key_pad;         // [B, Tk] or null
  const int32_t* kv_steps;        // [B] 64-key tiles up to the last valid key, or null (ldmatrix form only)
  bf *dq, *dk, *dv;
  int64_t lddq, lddk, lddv;
  int H, Tq, Tk;
  float scale;
};

// dsum[b, h, t] = sum_c dO[b*Tq + t, h*dh + c] * O[b*Tq + t, h*dh + c].  A CTA takes 32 consecutive rows, four per warp:
// the lanes sweep a row of both tensors in 16-byte chunks (512 contiguous bytes per load instruction, all loads of a row
// in flight together), the eight products of a chunk go to shared memory, one lane per head adds the head's dh / 8
// chunk sums, and the 32 x H block is written with consecutive threads on consecutive t (the output is [b, h, t]).
// The thread-per-(row, head) form issued its twelve 16-byte loads one after the other at a 192-byte lane stride and
// scattered its 4-byte results Tq apart: 0.5 ms per 256 k rows, a quarter of the HBM roofline.
constexpr int DSUM_ROWS = 32;
__global__ void __launch_bounds__(256)
attn_bwd_dsum_kernel(const bf* __restrict__ d_out, int64_t lddo, const bf* __restrict__ out, int64_t ldo,
                     float* __restrict__ dsum, int64_t rows, int H, int Tq, int dh) {
  extern __shared__ float ds_smem[];
  const int cpr = H * dh / 8, cph = dh / 8;          // 16-byte chunks per row / per head
  float* part = ds_smem;                             // [8 warps][cpr]
  float* hs = ds_smem + 8 * cpr;                     // [DSUM_ROWS][H + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_blocks = (rows + DSUM_ROWS - 1) / DSUM_ROWS;
  for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const int64_t base = blk * DSUM_ROWS;
#pragma unroll 1
    for (int k = 0; k < DSUM_ROWS / 8; ++k) {
      const int r = warp * (DSUM_ROWS / 8) + k;
      const int64_t row = base + r;
      if (row < rows) {
        const uint4* a4 = reinterpret_cast<const uint4*>(d_out + row * lddo);
        const uint4* o4 = reinterpret_cast<const uint4*>(out + row * ldo);
#pragma unroll 4
        for (int c = lane; c < cpr; c += 32) {
          const uint4 a = __ldg(a4 + c), o = __ldg(o4 + c);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
          float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 af = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[i]));
            const float2 of = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[i]));
            acc0 = fmaf(af.x, of.x, acc0);
            acc1 = fmaf(af.y, of.y, acc1);
          }
          part[warp * cpr + c] = acc0 + acc1;
        }
        __syncwarp();
        for (int h = lane; h < H; h += 32) {
          float s = 0.0f;
          for (int c = 0; c < cph; ++c) s += part[warp * cpr + h * cph + c];   // fixed order: deterministic
          hs[r * (H + 1) + h] = s;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < DSUM_ROWS * H; idx += 256) {
      const int h = idx / DSUM_ROWS, r = idx - h * DSUM_ROWS;
      const int64_t row = base + r;
      if (row < rows) {
        const int64_t b = row / Tq;
        dsum[(b * H + h) * Tq + (row - b * Tq)] = hs[r * (H + 1) + h];
      }
    }
    __syncthreads();
  }
}

// S and dP of one 64 x 64 tile pair, then P and dS of this thread's 16 elements.  Returns them in p[][] / ds[][].
template <int DH, bool USE_MMA>
__device__ __forceinline__ void tile_p_ds(const bf* sQ, const bf* sdO, const bf* sK, const bf* sV, const float* s_lse,
                                          const float* s_dsum, const float* s_kvalid, float scale, int wm, int wn, int lane,
                                          float (&p)[4][4], float (&ds)[4][4]) {
  constexpr int LDH = DH + 8;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) { p[nt][e] = 0.0f; ds[nt][e] = 0.0f; }
  warp_gemm<4, USE_MMA>(p, sQ + wm * 16 * LDH, LDH, sK + wn * 32 * LDH, LDH, DH, lane);     // S = Q K^T
  warp_gemm<4, USE_MMA>(ds, sdO + wm * 16 * LDH, LDH, sV + wn * 32 * LDH, LDH, DH, lane);   // dP = dO V^T
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = wm * 16 + g + (e >> 1) * 8;
      const int k = wn * 32 + nt * 8 + 2 * t + (e & 1);
      // rows past Tq carry lse = +inf (=> P = 0), PAD / out-of-range keys are switched off by s_kvalid
      const float pv = s_kvalid[k] != 0.0f ? __expf(p[nt][e] * scale - s_lse[q]) : 0.0f;
      p[nt][e] = pv;
      ds[nt][e] = pv * (ds[nt][e] - s_dsum[q]) * scale;
    }
}

template <int DH>
struct AttnBwdSmem {
  static constexpr int LDH = DH + 8;
  static constexpr int TILE = AB_T * LDH;        // elements of a [64][DH + 8] tile
  static constexpr int TILE_T = DH * AB_LDT;     // elements of a transposed [DH][72] tile
  static constexpr int SQ_T = AB_T * AB_LDT;     // elements of a [64][72] tile
  // dkv pass: K, V, Q, dO, Q^T, dO^T, P^T, dS^T ; dq pass: Q, dO, K, V, K^T, dS
  static constexpr size_t DKV_BYTES = (4 * TILE + 2 * TILE_T + 2 * SQ_T) * sizeof(bf) + 3 * AB_T * sizeof(float);
  static constexpr size_t DQ_BYTES = (4 * TILE + TILE_T + SQ_T) * sizeof(bf) + 3 * AB_T * sizeof(float);
};

// ---- pass 1: dK, dV of one 64-key tile of one (utterance, head)
template <int DH, bool USE_MMA>
__global__ void __launch_bounds__(AB_THREADS, 2)
attn_bwd_dkv_kernel(const AttnBwdParams p) {
  using SM = AttnBwdSmem<DH>;
  constexpr int NT = DH / 16;   // 8-column tiles per warp over half the head dim
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf* sK = reinterpret_cast<bf*>(smem_raw);
  bf* sV = sK + SM::TILE;
  bf* sQ = sV + SM::TILE;
  bf* sdO = sQ + SM::TILE;
  bf* sQT = sdO + SM::TILE;
  bf* sdOT = sQT + SM::TILE_T;
  bf* sPT = sdOT + SM::TILE_T;
  bf* sdST = sPT + SM::SQ_T;
  float* s_lse = reinterpret_cast<float*>(sdST + SM::SQ_T);
  float* s_dsum = s_lse + AB_T;
  float* s_kvalid = s_dsum + AB_T;

  const int k0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int64_t krow0 = static_cast<int64_t>(b) * p.Tk, qrow0 = static_cast<int64_t>(b) * p.Tq;

  load_tile<DH>(p.k + h * DH, p.ldk, krow0 + k0, krow0 + p.Tk, sK, nullptr);
  load_tile<DH>(p.v + h * DH, p.ldv, krow0 + k0, krow0 + p.Tk, sV, nullptr);
  if (threadIdx.x < AB_T) {
    const int k = k0 + threadIdx.x;
    s_kvalid[threadIdx.x] = (k < p.Tk && (p.key_pad == nullptr || p.key_pad[krow0 + k] == 0)) ? 1.0f : 0.0f;
  }
  float dk_acc[NT][4], dv_acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk_acc[nt][e] = 0.0f; dv_acc[nt][e] = 0.0f; }

  const float* lse = p.lse + (static_cast<int64_t>(b) * p.H + h) * p.Tq;
  const float* dsum = p.dsum + (static_cast<int64_t>(b) * p.H + h) * p.Tq;
  TileRegs<DH> rq, rdo;   // the next query tile, in flight while the current one is multiplied
  fetch_tile<DH>(rq, p.q + h * DH, p.ldq, qrow0, qrow0 + p.Tq);
  fetch_tile<DH>(rdo, p.d_out + h * DH, p.lddo, qrow0, qrow0 + p.Tq);
  for (int q0 = 0; q0 < p.Tq; q0 += AB_T) {
    __syncthreads();   // the previous tile pair's shared-memory reads are done
    stash_tile<DH>(rq, sQ, sQT);
    stash_tile<DH>(rdo, sdO, sdOT);
    if (threadIdx.x < AB_T) {
      const int q = q0 + threadIdx.x;
      s_lse[threadIdx.x] = q < p.Tq ? lse[q] : INFINITY;
      s_dsum[threadIdx.x] = q < p.Tq ? dsum[q] : 0.0f;
    }
    __syncthreads();
    if (q0 + AB_T < p.Tq) {
      fetch_tile<DH>(rq, p.q + h * DH, p.ldq, qrow0 + q0 + AB_T, qrow0 + p.Tq);
      fetch_tile<DH>(rdo, p.d_out + h * DH, p.lddo, qrow0 + q0 + AB_T, qrow0 + p.Tq);
    }
    float pr[4][4], ds[4][4];
    tile_p_ds<DH, USE_MMA>(sQ, sdO, sK, sV, s_lse, s_dsum, s_kvalid, p.scale, wm, wn, lane, pr, ds);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int q = wm * 16 + g + (e >> 1) * 8;
        const int k = wn * 32 + nt * 8 + 2 * t + (e & 1);
        sPT[k * AB_LDT + q] = __float2bfloat16_rn(pr[nt][e]);
        sdST[k * AB_LDT + q] = __float2bfloat16_rn(ds[nt][e]);
      }
    __syncthreads();
    // dV += P^T dO (rows = keys, contraction over the 64 queries), dK += dS^T Q
    warp_gemm<NT, USE_MMA>(dv_acc, sPT + wm * 16 * AB_LDT, AB_LDT, sdOT + wn * (DH / 2) * AB_LDT, AB_LDT, AB_T, lane);
    warp_gemm<NT, USE_MMA>(dk_acc, sdST + wm * 16 * AB_LDT, AB_LDT, sQT + wn * (DH / 2) * AB_LDT, AB_LDT, AB_T, lane);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k = k0 + wm * 16 + g + half * 8;
      if (k < p.Tk) {
        const int c = h * DH + wn * (DH / 2) + nt * 8 + 2 * t;
        *reinterpret_cast<__nv_bfloat162*>(p.dk + (krow0 + k) * p.lddk + c) =
            __floats2bfloat162_rn(dk_acc[nt][half * 2], dk_acc[nt][half * 2 + 1]);
        *reinterpret_cast<__nv_bfloat162*>(p.dv + (krow0 + k) * p.lddv + c) =
            __floats2bfloat162_rn(dv_acc[nt][half * 2], dv_acc[nt][half * 2 + 1]);
      }
    }
}

// ---- pass 2: dQ of one 64-query tile of one (utterance, head)
template <int DH, bool USE_MMA>
__global__ void __launch_bounds__(AB_THREADS, 2)
attn_bwd_dq_kernel(const AttnBwdParams p) {
  using SM = AttnBwdSmem<DH>;
  constexpr int NT = DH / 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf* sQ = reinterpret_cast<bf*>(smem_raw);
  bf* sdO = sQ + SM::TILE;
  bf* sK = sdO + SM::TILE;
  bf* sV = sK + SM::TILE;
  bf* sKT = sV + SM::TILE;
  bf* sdS = sKT + SM::TILE_T;
  float* s_lse = reinterpret_cast<float*>(sdS + SM::SQ_T);
  float* s_dsum = s_lse + AB_T;
  float* s_kvalid = s_dsum + AB_T;

  const int q0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int64_t krow0 = static_cast<int64_t>(b) * p.Tk, qrow0 = static_cast<int64_t>(b) * p.Tq;

  load_tile<DH>(p.q + h * DH, p.ldq, qrow0 + q0, qrow0 + p.Tq, sQ, nullptr);
  load_tile<DH>(p.d_out + h * DH, p.lddo, qrow0 + q0, qrow0 + p.Tq, sdO, nullptr);
  if (threadIdx.x < AB_T) {
    const int q = q0 + threadIdx.x;
    const int64_t o = (static_cast<int64_t>(b) * p.H + h) * p.Tq + q;
    s_lse[threadIdx.x] = q < p.Tq ? p.lse[o] : INFINITY;
    s_dsum[threadIdx.x] = q < p.Tq ? p.dsum[o] : 0.0f;
  }
  float dq_acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq_acc[nt][e] = 0.0f;

  TileRegs<DH> rk, rv;   // the next key tile, in flight while the current one is multiplied
  fetch_tile<DH>(rk, p.k + h * DH, p.ldk, krow0, krow0 + p.Tk);
  fetch_tile<DH>(rv, p.v + h * DH, p.ldv, krow0, krow0 + p.Tk);
  for (int k0 = 0; k0 < p.Tk; k0 += AB_T) {
    __syncthreads();
    stash_tile<DH>(rk, sK, sKT);
    stash_tile<DH>(rv, sV, nullptr);
    if (threadIdx.x < AB_T) {
      const int k = k0 + threadIdx.x;
      s_kvalid[threadIdx.x] = (k < p.Tk && (p.key_pad == nullptr || p.key_pad[krow0 + k] == 0)) ? 1.0f : 0.0f;
    }
    __syncthreads();
    if (k0 + AB_T < p.Tk) {
      fetch_tile<DH>(rk, p.k + h * DH, p.ldk, krow0 + k0 + AB_T, krow0 + p.Tk);
      fetch_tile<DH>(rv, p.v + h * DH, p.ldv, krow0 + k0 + AB_T, krow0 + p.Tk);
    }
    float pr[4][4], ds[4][4];
    tile_p_ds<DH, USE_MMA>(sQ, sdO, sK, sV, s_lse, s_dsum, s_kvalid, p.scale, wm, wn, lane, pr, ds);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int q = wm * 16 + g + half * 8;
        const int k = wn * 32 + nt * 8 + 2 * t;
        *reinterpret_cast<__nv_bfloat162*>(sdS + q * AB_LDT + k) = __floats2bfloat162_rn(ds[nt][half * 2], ds[nt][half * 2 + 1]);
      }
    __syncthreads();
    // dQ += dS K (rows = queries, contraction over the 64 keys)
    warp_gemm<NT, USE_MMA>(dq_acc, sdS + wm * 16 * AB_LDT, AB_LDT, sKT + wn * (DH / 2) * AB_LDT, AB_LDT, AB_T, lane);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int q = q0 + wm * 16 + g + half * 8;
      if (q < p.Tq) {
        const int c = h * DH + wn * (DH / 2) + nt * 8 + 2 * t;
        *reinterpret_cast<__nv_bfloat162*>(p.dq + (qrow0 + q) * p.lddq + c) =
            __floats2bfloat162_rn(dq_acc[nt][half * 2], dq_acc[nt][half * 2 + 1]);
      }
    }
}

template <int DH, bool USE_MMA>
static int launch_attn_bwd(const AttnBwdParams& p, int B, cudaStream_t s) {
  using SM = AttnBwdSmem<DH>;
  static uint64_t attr_done = 0;
  if (device_needs_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<DH, USE_MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(SM::DKV_BYTES));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dq_kernel<DH, USE_MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(SM::DQ_BYTES));
    if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "attention_backward: %s", cudaGetErrorString(e));
  }
  attn_bwd_dkv_kernel<DH, USE_MMA><<<dim3((p.Tk + AB_T - 1) / AB_T, p.H, B), AB_THREADS, SM::DKV_BYTES, s>>>(p);
  int rc = check_launch("attention_backward (dK, dV)");
  if (rc) return rc;
  attn_bwd_dq_kernel<DH, USE_MMA><<<dim3((p.Tq + AB_T - 1) / AB_T, p.H, B), AB_THREADS, SM::DQ_BYTES, s>>>(p);
  return check_launch("attention_backward (dQ)");
}


// =====================================================================================================================
// ldmatrix form (the default): the same two passes, but every fragment comes from ldmatrix(.trans), so NO transposed
// copy of any tile exists in shared memory: the contraction-major operands of dV = P^T dO, dK = dS^T Q and dQ = dS K
// are read with ldmatrix.trans straight from the row-major tiles.  Global loads are coalesced (lanes along the columns
// of a row); P and dS are written [query][key] with 4-byte stores.  ncu of the first form
// (profiles/r01_ncu_attention_bwd_v14_summary.txt): the LSU data pipe at 49 - 73 % of peak, 38 wavefronts per global
// request from the row-per-lane loads that the transposed stores had asked for, mio_throttle the top stall.
// =====================================================================================================================
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const bf* ptr) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const bf* ptr) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// acc[16 x NT*8] += A . B^T over K (multiple of 16).  The 8x8 matrices of an x4 load are addressed by lane groups of 8
// (mi = lane / 8); thread (g, t) receives (row g, elements 2t, 2t+1) of each, or of its transpose with .trans.
//   A_TRANS = false: A stored [m][k] (row pitch lda), A points at (m0, 0).
//   A_TRANS = true : A stored [k][m], A points at (0, m0)   (A = S^T: dV = P^T dO, dK = dS^T Q).
//   B_TRANS = false: B stored [n][k] (row pitch ldb), B points at (n0, 0).
//   B_TRANS = true : B stored [k][n], B points at (0, n0)   (dO, Q, K read as they lie).
template <int NT, bool A_TRANS, bool B_TRANS>
__device__ __forceinline__ void warp_gemm_lm(float (&acc)[NT][4], const bf* __restrict__ A, int lda,
                                             const bf* __restrict__ B, int ldb, int K, int lane) {
  static_assert(NT % 2 == 0, "B fragments are loaded two 8-column tiles at a time");
  const int r8 = lane & 7, lo = (lane >> 3) & 1, hi = lane >> 4;
  for (int k0 = 0; k0 < K; k0 += 16) {
    uint32_t a0, a1, a2, a3;
    if constexpr (!A_TRANS) {
      // matrices: (m 0-7, k 0-7) (m 8-15, k 0-7) (m 0-7, k 8-15) (m 8-15, k 8-15)
      ldsm_x4(a0, a1, a2, a3, A + (lo * 8 + r8) * lda + k0 + hi * 8);
    } else {
      // source blocks: (k 0-7, m 0-7) (k 0-7, m 8-15) (k 8-15, m 0-7) (k 8-15, m 8-15), transposed on the way
      ldsm_x4_trans(a0, a1, a2, a3, A + (k0 + hi * 8 + r8) * lda + lo * 8);
    }
#pragma unroll
    for (int nt = 0; nt < NT; nt += 2) {
      uint32_t b0, b1, b2, b3;   // (b0, b1) of column tile nt, (b0, b1) of column tile nt + 1
      if constexpr (!B_TRANS) {
        // matrices: (n 0-7, k 0-7) (n 0-7, k 8-15) (n 8-15, k 0-7) (n 8-15, k 8-15)
        ldsm_x4(b0, b1, b2, b3, B + (nt * 8 + hi * 8 + r8) * ldb + k0 + lo * 8);
      } else {
        // source blocks: (k 0-7, n 0-7) (k 8-15, n 0-7) (k 0-7, n 8-15) (k 8-15, n 8-15), transposed on the way
        ldsm_x4_trans(b0, b1, b2, b3, B + (k0 + lo * 8 + r8) * ldb + nt * 8 + hi * 8);
      }
      mma_16816(acc[nt], a0, a1, a2, a3, b0, b1);
      mma_16816(acc[nt + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

// Coalesced tile staging: consecutive lanes take consecutive 16-byte chunks of a row.
template <int DH>
__device__ __forceinline__ void fetch_tile_rows(TileRegs<DH>& regs, const bf* __restrict__ src, int64_t ld, int64_t row0,
                                                int64_t row_end) {
  constexpr int CH = DH / 8;
#pragma unroll
  for (int j = 0; j < TileRegs<DH>::N; ++j) {
    const int idx = threadIdx.x + j * AB_THREADS;
    const int r = idx / CH, c = (idx - r * CH) * 8;
    regs.v[j] = make_uint4(0u, 0u, 0u, 0u);
    if (row0 + r < row_end) regs.v[j] = __ldg(reinterpret_cast<const uint4*>(src + (row0 + r) * ld + c));
  }
}

template <int DH>
__device__ __forceinline__ void stash_tile_rows(const TileRegs<DH>& regs, bf* __restrict__ dst) {
  constexpr int LDH = DH + 8, CH = DH / 8;
#pragma unroll
  for (int j = 0; j < TileRegs<DH>::N; ++j) {
    const int idx = threadIdx.x + j * AB_THREADS;
    const int r = idx / CH, c = (idx - r * CH) * 8;
    *reinterpret_cast<uint4*>(dst + r * LDH + c) = regs.v[j];
  }
}

// S and dP of one tile pair -> P and dS of this thread's 16 elements, written [query][key] into sP / sdS (either may be
// null when the pass does not need it).
template <int DH>
__device__ __forceinline__ void tile_p_ds_lm(const bf* sQ, const bf* sdO, const bf* sK, const bf* sV, const float* s_lse,
                                             const float* s_dsum, const float* s_kvalid, float scale, int wm, int wn,
                                             int lane, bf* sP, bf* sdS) {
  constexpr int LDH = DH + 8;
  float p[4][4], ds[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) { p[nt][e] = 0.0f; ds[nt][e] = 0.0f; }
  warp_gemm_lm<4, false, false>(p, sQ + wm * 16 * LDH, LDH, sK + wn * 32 * LDH, LDH, DH, lane);     // S = Q K^T
  warp_gemm_lm<4, false, false>(ds, sdO + wm * 16 * LDH, LDH, sV + wn * 32 * LDH, LDH, DH, lane);   // dP = dO V^T
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int q = wm * 16 + g + half * 8;
      const int k = wn * 32 + nt * 8 + 2 * t;
      const float lse = s_lse[q], dsum = s_dsum[q];
      const float p0 = s_kvalid[k] != 0.0f ? __expf(p[nt][half * 2] * scale - lse) : 0.0f;
      const float p1 = s_kvalid[k + 1] != 0.0f ? __expf(p[nt][half * 2 + 1] * scale - lse) : 0.0f;
      if (sP != nullptr) *reinterpret_cast<__nv_bfloat162*>(sP + q * AB_LDT + k) = __floats2bfloat162_rn(p0, p1);
      *reinterpret_cast<__nv_bfloat162*>(sdS + q * AB_LDT + k) =
          __floats2bfloat162_rn(p0 * (ds[nt][half * 2] - dsum) * scale, p1 * (ds[nt][half * 2 + 1] - dsum) * scale);
    }
}

template <int DH>
struct AttnBwdSmemLm {
  static constexpr int TILE = AB_T * (DH + 8);
  static constexpr int SQ_T = AB_T * AB_LDT;
  static constexpr size_t DKV_BYTES = (4 * TILE + 2 * SQ_T) * sizeof(bf) + 3 * AB_T * sizeof(float);   // K V Q dO P dS
  static constexpr size_t DQ_BYTES = (4 * TILE + SQ_T) * sizeof(bf) + 3 * AB_T * sizeof(float);        // Q dO K V dS
};

template <int DH>
__global__ void __launch_bounds__(AB_THREADS, 2)
attn_bwd_dkv_lm_kernel(const AttnBwdParams p) {
  using SM = AttnBwdSmemLm<DH>;
  constexpr int NT = DH / 16, LDH = DH + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf* sK = reinterpret_cast<bf*>(smem_raw);
  bf* sV = sK + SM::TILE;
  bf* sQ = sV + SM::TILE;
  bf* sdO = sQ + SM::TILE;
  bf* sP = sdO + SM::TILE;
  bf* sdS = sP + SM::SQ_T;
  float* s_lse = reinterpret_cast<float*>(sdS + SM::SQ_T);
  float* s_dsum = s_lse + AB_T;
  float* s_kvalid = s_dsum + AB_T;

  const int k0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int64_t krow0 = static_cast<int64_t>(b) * p.Tk, qrow0 = static_cast<int64_t>(b) * p.Tq;

  if (p.kv_steps != nullptr && static_cast<int>(blockIdx.x) >= p.kv_steps[b]) {
    // a tile of trailing PAD keys: P = 0 for every query, so dK = dV = 0 without looking at anything
    constexpr int CH = DH / 8;
    for (int idx = threadIdx.x; idx < AB_T * CH; idx += AB_THREADS) {
      const int r = idx / CH, c = (idx - r * CH) * 8;
      if (k0 + r < p.Tk) {
        *reinterpret_cast<uint4*>(p.dk + (krow0 + k0 + r) * p.lddk + h * DH + c) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(p.dv + (krow0 + k0 + r) * p.lddv + h * DH + c) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    return;
  }
  TileRegs<DH> rq, rdo;
  fetch_tile_rows<DH>(rq, p.k + h * DH, p.ldk, krow0 + k0, krow0 + p.Tk);
  fetch_tile_rows<DH>(rdo, p.v + h * DH, p.ldv, krow0 + k0, krow0 + p.Tk);
  stash_tile_rows<DH>(rq, sK);
  stash_tile_rows<DH>(rdo, sV);
  if (threadIdx.x < AB_T) {
    const int k = k0 + threadIdx.x;
    s_kvalid[threadIdx.x] = (k < p.Tk && (p.key_pad == nullptr || p.key_pad[krow0 + k] == 0)) ? 1.0f : 0.0f;
  }
  float dk_acc[NT][4], dv_acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk_acc[nt][e] = 0.0f; dv_acc[nt][e] = 0.0f; }

  const float* lse = p.lse + (static_cast<int64_t>(b) * p.H + h) * p.Tq;
  const float* dsum = p.dsum + (static_cast<int64_t>(b) * p.H + h) * p.Tq;
  fetch_tile_rows<DH>(rq, p.q + h * DH, p.ldq, qrow0, qrow0 + p.Tq);
  fetch_tile_rows<DH>(rdo, p.d_out + h * DH, p.lddo, qrow0, qrow0 + p.Tq);
  for (int q0 = 0; q0 < p.Tq; q0 += AB_T) {
    __syncthreads();   // the previous tile pair's shared-memory reads are done
    stash_tile_rows<DH>(rq, sQ);
    stash_tile_rows<DH>(rdo, sdO);
    if (threadIdx.x < AB_T) {
      const int q = q0 + threadIdx.x;
      s_lse[threadIdx.x] = q < p.Tq ? lse[q] : INFINITY;
      s_dsum[threadIdx.x] = q < p.Tq ? dsum[q] : 0.0f;
    }
    __syncthreads();
    if (q0 + AB_T < p.Tq) {
      fetch_tile_rows<DH>(rq, p.q + h * DH, p.ldq, qrow0 + q0 + AB_T, qrow0 + p.Tq);
      fetch_tile_rows<DH>(rdo, p.d_out + h * DH, p.lddo, qrow0 + q0 + AB_T, qrow0 + p.Tq);
    }
    tile_p_ds_lm<DH>(sQ, sdO, sK, sV, s_lse, s_dsum, s_kvalid, p.scale, wm, wn, lane, sP, sdS);
    __syncthreads();
    // dV[k][:] += sum_q P[q][k] dO[q][:],  dK[k][:] += sum_q dS[q][k] Q[q][:]   (rows = this warp's 16 keys)
    warp_gemm_lm<NT, true, true>(dv_acc, sP + wm * 16, AB_LDT, sdO + wn * (DH / 2), LDH, AB_T, lane);
    warp_gemm_lm<NT, true, true>(dk_acc, sdS + wm * 16, AB_LDT, sQ + wn * (DH / 2), LDH, AB_T, lane);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k = k0 + wm * 16 + g + half * 8;
      if (k < p.Tk) {
        const int c = h * DH + wn * (DH / 2) + nt * 8 + 2 * t;
        *reinterpret_cast<__nv_bfloat162*>(p.dk + (krow0 + k) * p.lddk + c) =
            __floats2bfloat162_rn(dk_acc[nt][half * 2], dk_acc[nt][half * 2 + 1]);
        *reinterpret_cast<__nv_bfloat162*>(p.dv + (krow0 + k) * p.lddv + c) =
            __floats2bfloat162_rn(dv_acc[nt][half * 2], dv_acc[nt][half * 2 + 1]);
      }
    }
}

template <int DH>
__global__ void __launch_bounds__(AB_THREADS, 2)
attn_bwd_dq_lm_kernel(const AttnBwdParams p) {
  using SM = AttnBwdSmemLm<DH>;
  constexpr int NT = DH / 16, LDH = DH + 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf* sQ = reinterpret_cast<bf*>(smem_raw);
  bf* sdO = sQ + SM::TILE;
  bf* sK = sdO + SM::TILE;
  bf* sV = sK + SM::TILE;
  bf* sdS = sV + SM::TILE;
  float* s_lse = reinterpret_cast<float*>(sdS + SM::SQ_T);
  float* s_dsum = s_lse + AB_T;
  float* s_kvalid = s_dsum + AB_T;

  const int q0 = blockIdx.x * AB_T, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int64_t krow0 = static_cast<int64_t>(b) * p.Tk, qrow0 = static_cast<int64_t>(b) * p.Tq;

  TileRegs<DH> rk, rv;
  fetch_tile_rows<DH>(rk, p.q + h * DH, p.ldq, qrow0 + q0, qrow0 + p.Tq);
  fetch_tile_rows<DH>(rv, p.d_out + h * DH, p.lddo, qrow0 + q0, qrow0 + p.Tq);
  stash_tile_rows<DH>(rk, sQ);
  stash_tile_rows<DH>(rv, sdO);
  if (threadIdx.x < AB_T) {
    const int q = q0 + threadIdx.x;
    const int64_t o = (static_cast<int64_t>(b) * p.H + h) * p.Tq + q;
    s_lse[threadIdx.x] = q < p.Tq ? p.lse[o] : INFINITY;
    s_dsum[threadIdx.x] = q < p.Tq ? p.dsum[o] : 0.0f;
  }
  float dq_acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq_acc[nt][e] = 0.0f;

  // trailing tiles of PAD keys contribute P = 0: not visited
  const int k_end = p.kv_steps != nullptr ? min(p.Tk, p.kv_steps[b] * AB_T) : p.Tk;
  fetch_tile_rows<DH>(rk, p.k + h * DH, p.ldk, krow0, krow0 + p.Tk);
  fetch_tile_rows<DH>(rv, p.v + h * DH, p.ldv, krow0, krow0 + p.Tk);
  for (int k0 = 0; k0 < k_end; k0 += AB_T) {
    __syncthreads();
    stash_tile_rows<DH>(rk, sK);
    stash_tile_rows<DH>(rv, sV);
    if (threadIdx.x < AB_T) {
      const int k = k0 + threadIdx.x;
      s_kvalid[threadIdx.x] = (k < p.Tk && (p.key_pad == nullptr || p.key_pad[krow0 + k] == 0)) ? 1.0f : 0.0f;
    }
    __syncthreads();
    if (k0 + AB_T < k_end) {
      fetch_tile_rows<DH>(rk, p.k + h * DH, p.ldk, krow0 + k0 + AB_T, krow0 + p.Tk);
      fetch_tile_rows<DH>(rv, p.v + h * DH, p.ldv, krow0 + k0 + AB_T, krow0 + p.Tk);
    }
    tile_p_ds_lm<DH>(sQ, sdO, sK, sV, s_lse, s_dsum, s_kvalid, p.scale, wm, wn, lane, nullptr, sdS);
    __syncthreads();
    // dQ[q][:] += sum_k dS[q][k] K[k][:]   (rows = this warp's 16 queries; K read as it lies)
    warp_gemm_lm<NT, false, true>(dq_acc, sdS + wm * 16 * AB_LDT, AB_LDT, sK + wn * (DH / 2), LDH, AB_T, lane);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int q = q0 + wm * 16 + g + half * 8;
      if (q < p.Tq) {
        const int c = h * DH + wn * (DH / 2) + nt * 8 + 2 * t;
        *reinterpret_cast<__nv_bfloat162*>(p.dq + (qrow0 + q) * p.lddq + c) =
            __floats2bfloat162_rn(dq_acc[nt][half * 2], dq_acc[nt][half * 2 + 1]);
      }
    }
}

template <int DH>
static int launch_attn_bwd_lm(const AttnBwdParams& p, int B, cudaStream_t s) {
  using SM = AttnBwdSmemLm<DH>;
  static uint64_t attr_done = 0;
  if (device_needs_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkv_lm_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(SM::DKV_BYTES));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dq_lm_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(SM::DQ_BYTES));
    if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "attention_backward: %s", cudaGetErrorString(e));
  }
  attn_bwd_dkv_lm_kernel<DH><<<dim3((p.Tk + AB_T - 1) / AB_T, p.H, B), AB_THREADS, SM::DKV_BYTES, s>>>(p);
  int rc = check_launch("attention_backward (dK, dV)");
  if (rc) return rc;
  attn_bwd_dq_lm_kernel<DH><<<dim3((p.Tq + AB_T - 1) / AB_T, p.H, B), AB_THREADS, SM::DQ_BYTES, s>>>(p);
  return check_launch("attention_backward (dQ)");
}

}  // namespace hriemo

using namespace hriemo;

namespace hriemo {
template <int DH>
int launch_attn_bwd_tc(const hriemo_attn_bwd_args& a, cudaStream_t stream);   // attention_bwd_tc.cu
}

extern "C" int hriemo_attention_backward_bf16(const hriemo_attn_bwd_args* a, void* stream) {
  HRIEMO_REQUIRE(a != nullptr, "attention_backward: null args");
  HRIEMO_REQUIRE(a->q && a->k && a->v && a->out && a->d_out && a->lse && a->dsum && a->dq && a->dk && a->dv,
                 "attention_backward: null pointer");
  HRIEMO_REQUIRE(a->B > 0 && a->B <= 65535 && a->H > 0 && a->H <= 65535 && a->Tq > 0 && a->Tk > 0,
                 "attention_backward: bad shape B=%d H=%d Tq=%d Tk=%d", a->B, a->H, a->Tq, a->Tk);
  HRIEMO_REQUIRE(a->impl >= 0 && a->impl <= 4,
                 "attention_backward: impl=%d (0 / 3 tcgen05 form, 1 FMA, 2 first mma.sync form, 4 ldmatrix mma.sync form)", a->impl);
  HRIEMO_REQUIRE(a->kv_steps == nullptr || a->key_pad != nullptr, "attention_backward: kv_steps comes with key_pad");
  HRIEMO_REQUIRE(a->drop_p8 == 0u || ((a->impl == 0 || a->impl == 3) && a->drop_p8 <= 255u && a->drop_scale > 0.0f),
                 "attention_backward: dropout on the probabilities needs the tcgen05 form (impl 0) and drop_scale = 1 / (1 - p)");
  HRIEMO_REQUIRE(a->dh == 32 || a->dh == 64 || a->dh == 96 || a->dh == 128, "attention_backward: dh=%d not in {32, 64, 96, 128}",
                 a->dh);
  const int64_t lds[] = {a->ldq, a->ldk, a->ldv, a->ldo, a->lddo, a->lddq, a->lddk, a->lddv};
  for (int64_t ld : lds) HRIEMO_REQUIRE(ld % 8 == 0 && ld >= static_cast<int64_t>(a->H) * a->dh, "attention_backward: bad leading dimension");
  const void* ptrs[] = {a->q, a->k, a->v, a->out, a->d_out, a->dq, a->dk, a->dv};
  for (const void* ptr : ptrs)
    HRIEMO_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "attention_backward: operands must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t rows = static_cast<int64_t>(a->B) * a->Tq;
  int64_t gb = (rows + DSUM_ROWS - 1) / DSUM_ROWS;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;
  if (gb > cap) gb = cap;
  const size_t ds_bytes = sizeof(float) * (static_cast<size_t>(8) * (a->H * a->dh / 8) + static_cast<size_t>(DSUM_ROWS) * (a->H + 1));
  HRIEMO_REQUIRE(ds_bytes <= 48 * 1024, "attention_backward: H * dh = %d too wide for the D kernel", a->H * a->dh);
  attn_bwd_dsum_kernel<<<static_cast<unsigned>(gb), 256, ds_bytes, s>>>(static_cast<const bf*>(a->d_out), a->lddo,
                                                               static_cast<const bf*>(a->out), a->ldo, a->dsum, rows, a->H,
                                                               a->Tq, a->dh);
  int rc = check_launch("attention_backward (D)");
  if (rc) return rc;
  if (a->impl == 0 || a->impl == 3) {   // the tcgen05 form (attention_bwd_tc.cu)
    switch (a->dh) {
      case 32: return launch_attn_bwd_tc<32>(*a, s);
      case 64: return launch_attn_bwd_tc<64>(*a, s);
      case 96: return launch_attn_bwd_tc<96>(*a, s);
      default: return launch_attn_bwd_tc<128>(*a, s);
    }
  }
  AttnBwdParams p;
  p.q = static_cast<const bf*>(a->q); p.k = static_cast<const bf*>(a->k); p.v = static_cast<const bf*>(a->v);
  p.d_out = static_cast<const bf*>(a->d_out);
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.lddo = a->lddo;
  p.lse = a->lse; p.dsum = a->dsum; p.key_pad = a->key_pad; p.kv_steps = a->kv_steps;
  p.dq = static_cast<bf*>(a->dq); p.dk = static_cast<bf*>(a->dk); p.dv = static_cast<bf*>(a->dv);
  p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  p.H = a->H; p.Tq = a->Tq; p.Tk = a->Tk; p.scale = a->scale;
#define HRIEMO_AB_DISPATCH(DHV)                                                         \
  case DHV:                                                                             \
    return a->impl == 1 ? launch_attn_bwd<DHV, false>(p, a->B, s)                    \
         : a->impl == 2 ? launch_attn_bwd<DHV, true>(p, a->B, s)                     \
                           : launch_attn_bwd_lm<DHV>(p, a->B, s)
  switch (a->dh) {
    HRIEMO_AB_DISPATCH(32);
    HRIEMO_AB_DISPATCH(64);
    HRIEMO_AB_DISPATCH(96);
    default:
      HRIEMO_AB_DISPATCH(128);
  }
#undef HRIEMO_AB_DISPATCH
}
