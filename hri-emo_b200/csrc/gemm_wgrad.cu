// Weight / bias gradient of y = x W^T + b on sm_100a tensor cores (the bulk of the
// backward FLOPs):   dW[N,K] = dY[M,N]^T . X[M,K]     db[N] = sum_m dY[m, n]
// (what loss.backward() accumulates into nn.Linear.weight.grad / .bias.grad at
// scripts/fusion/train_fusion_seq_level_decoder.py:332 for every projection and FFN layer of models/*.py).
//
// The contraction runs over the ROWS of both operands (M = B*T, about a million), so neither operand is
// K-major: both are consumed exactly as they lie in HBM, as MN-major UMMA operands -- the form the attention
// kernel uses for V -- instead of being transposed first.  One CTA owns a 128 (n) x BNK (k) tile of dW and one
// of `splits` ranges of m:
//   warp 0      TMA producer: per 64-row block of m, 4 boxes of dY [64 m][32 n] and BNK/32 boxes of X
//               [64 m][32 k], 64-byte-swizzled, into a 4-stage ring;
//   warp 1      tcgen05.mma issuer (one thread): D[128 x BNK] += A[128 x 16] B[16 x BNK], four per block, fp32
//               accumulation in tensor memory for the whole m range;
//   warps 2..5  epilogue: tcgen05.ld, fp32 partial tile to workspace[split][n][k].
// A second kernel adds the `splits` partial tiles in a fixed order (deterministic, unlike atomics) into dW,
// optionally on top of what is there (gradient accumulation).  db is a two-stage column sum of dY.
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace hriemo {

constexpr int WG_BN = 128;      // n rows of dW per tile (UMMA M)
constexpr int WG_MB = 64;       // rows of m per pipeline stage (four K=16 MMA steps)
constexpr int WG_STAGES = 4;
constexpr int WG_GROUP = WG_MB * 64;   // one [64 m][32 columns] box: 64 B per row, 64B swizzle
constexpr int WG_THREADS = 192;
constexpr uint32_t kUmmaAMajorMN = 1u << 15;   // instruction-descriptor flag: A operand is MN-major

template <int BNK>
struct WgradSmem {
  static constexpr int A_STAGE = (WG_BN / 32) * WG_GROUP;   // 16 KB
  static constexpr int B_STAGE = (BNK / 32) * WG_GROUP;     // 32 KB (BNK = 256) / 16 KB (128)
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = WG_STAGES * A_STAGE;
  static constexpr int BAR_OFF = B_OFF + WG_STAGES * B_STAGE;   // full[4] empty[4] tfull
  static constexpr int TMEM_SLOT_OFF = BAR_OFF + (2 * WG_STAGES + 1) * 8;
  static constexpr int DYN_BYTES = TMEM_SLOT_OFF + 16 + 1024;
};

template <int BNK>
__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                  float* __restrict__ partials, int64_t M, int N, int K, int blocks_per_split) {
  using L = WgradSmem<BNK>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base + L::A_OFF, sB = base + L::B_OFF;
  const uint32_t bar_full = base + L::BAR_OFF, bar_empty = bar_full + WG_STAGES * 8, bar_tfull = bar_empty + WG_STAGES * 8;
  const uint32_t tmem_slot = base + L::TMEM_SLOT_OFF;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * BNK, n0 = blockIdx.y * WG_BN, split = blockIdx.z;
  const int total_blocks = static_cast<int>((M + WG_MB - 1) / WG_MB);
  const int blk0 = split * blocks_per_split;
  const int blk1 = min(total_blocks, blk0 + blocks_per_split);
  const int nblk = max(0, blk1 - blk0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_dy);
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
    }
    mbar_init(bar_tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BNK>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nblk; ++i) {
        const uint32_t s = i % WG_STAGES, par = (i / WG_STAGES) & 1u;
        mbar_wait(bar_empty + s * 8, par ^ 1);
        mbar_arrive_expect_tx(bar_full + s * 8, L::A_STAGE + L::B_STAGE);
        const int m = (blk0 + i) * WG_MB;   // rows past M read as zero (no contribution)
        for (int c = 0; c < WG_BN / 32; ++c)
          tma_load_2d(&tm_dy, bar_full + s * 8, sA + s * L::A_STAGE + c * WG_GROUP, n0 + c * 32, m);
        for (int c = 0; c < BNK / 32; ++c)
          tma_load_2d(&tm_x, bar_full + s * 8, sB + s * L::B_STAGE + c * WG_GROUP, k0 + c * 32, m);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(WG_BN, BNK) | kUmmaAMajorMN | kUmmaBMajorMN;
      for (int i = 0; i < nblk; ++i) {
        const uint32_t s = i % WG_STAGES, par = (i / WG_STAGES) & 1u;
        mbar_wait(bar_full + s * 8, par);
        tc_fence_after_sync();
        const uint64_t a_desc = umma_desc_mn_sw64(sA + s * L::A_STAGE, WG_GROUP);
        const uint64_t b_desc = umma_desc_mn_sw64(sB + s * L::B_STAGE, WG_GROUP);
#pragma unroll
        for (int st = 0; st < WG_MB / 16; ++st)   // 16 rows of m = 16 x 64 B inside every column group
          umma_bf16(tmem_base, a_desc + ((st * 16 * 64) >> 4), b_desc + ((st * 16 * 64) >> 4), idesc, (i | st) != 0);
        umma_commit(bar_empty + s * 8);
      }
      if (nblk > 0) umma_commit(bar_tfull);
      else mbar_arrive(bar_tfull);   // an empty split still hands a (zero) tile to the epilogue
    }
  } else {
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

// partial[split][n] = sum over the split's rows of dY[m, n].  A warp reads 256 consecutive columns of a row (one
// 16-byte load per lane), the 8 warps of the block stride over the rows; fixed-order combine through shared memory.
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ dy, int64_t ld, int64_t M, int N, int64_t rows_per_split,
                      float* __restrict__ partial) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int64_t m0 = blockIdx.y * rows_per_split, m1 = min(M, m0 + rows_per_split);
  float acc[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  if (col < N) {   // N is a multiple of 128 and col of 8: a lane's eight columns are all in range or all out
    for (int64_t m = m0 + w; m < m1; m += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + m * ld + col));
      acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
      acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z); acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[w][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float s = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][threadIdx.x];
    partial[static_cast<int64_t>(blockIdx.y) * N + c] = s;
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
static int wgrad_splits(int64_t M, int N, int K, int bnk) {
  const int64_t tiles = static_cast<int64_t>(N / WG_BN) * (K / bnk);
  const int64_t blocks = (M + WG_MB - 1) / WG_MB;
  const int64_t sms = device_sm_count();
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
constexpr int WG_COLSUM_SPLITS = 64;

template <int BNK>
static int launch_wgrad(const void* dY, int64_t lddy, const void* X, int64_t ldx, int64_t M, int N, int K,
                        float* partials, int splits, cudaStream_t stream) {
  using L = WgradSmem<BNK>;
  CUtensorMap tm_dy, tm_x;
  int rc = make_tmap_bf16_2d(&tm_dy, dY, (uint64_t)N, (uint64_t)M, (uint64_t)lddy, 32, WG_MB, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_x, X, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, 32, WG_MB, 64);
  if (rc) return rc;
  static uint64_t attr_done = 0;
  if (device_needs_attr(&attr_done)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel<BNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return set_error(HRIEMO_ERR_CUDA, "linear_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const int64_t blocks = (M + WG_MB - 1) / WG_MB;
  const int bps = static_cast<int>((blocks + splits - 1) / splits);
  dim3 grid(K / BNK, N / WG_BN, splits);
  gemm_wgrad_kernel<BNK><<<grid, WG_THREADS, L::DYN_BYTES, stream>>>(tm_dy, tm_x, partials, M, N, K, bps);
  return check_launch("linear_wgrad");
}

}  // namespace hriemo

using namespace hriemo;

extern "C" int64_t hriemo_linear_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K) {
  if (M <= 0 || N <= 0 || K <= 0 || N % WG_BN != 0 || K % 128 != 0) return 0;
  const int splits = wgrad_splits(M, N, K, K % 256 == 0 ? 256 : 128);
  return (static_cast<int64_t>(splits) * N * K + static_cast<int64_t>(WG_COLSUM_SPLITS) * N) * static_cast<int64_t>(sizeof(float));
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
  int rc = bnk == 256 ? launch_wgrad<256>(dY, lddy, X, ldx, M, N, K, partials, splits, s)
                      : launch_wgrad<128>(dY, lddy, X, ldx, M, N, K, partials, splits, s);
  if (rc) return rc;
  const int64_t nk = static_cast<int64_t>(N) * K;
  int64_t gb = (nk + 255) / 256;
  if (gb > static_cast<int64_t>(device_sm_count()) * 8) gb = static_cast<int64_t>(device_sm_count()) * 8;
  const unsigned g = static_cast<unsigned>(gb);
  sum_partials_kernel<<<g, 256, 0, s>>>(partials, splits, nk, dW, accumulate);
  rc = check_launch("linear_wgrad (reduce)");
  if (rc || db == nullptr) return rc;
  float* bpart = partials + static_cast<int64_t>(splits) * nk;
  const int64_t rps = (M + WG_COLSUM_SPLITS - 1) / WG_COLSUM_SPLITS;
  colsum_partial_kernel<<<dim3((N + 255) / 256, WG_COLSUM_SPLITS), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(dY), lddy, M, N,
                                                                              rps, bpart);
  rc = check_launch("linear_wgrad (bias partials)");
  if (rc) return rc;
  sum_partials_kernel<<<(N + 255) / 256, 256, 0, s>>>(bpart, WG_COLSUM_SPLITS, N, db, accumulate);
  return check_launch("linear_wgrad (bias reduce)");
}

extern "C" int hriemo_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int32_t rows, int32_t cols,
                                     void* stream) {
  HRIEMO_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "transpose: bad argument");
  transpose_bf16_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), ld_in, static_cast<__nv_bfloat16*>(out), ld_out, rows, cols);
  return check_launch("transpose_bf16");
}
