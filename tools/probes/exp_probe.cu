// Micro-probe: how fast can one softmax warp turn 64 scores per thread into packed bf16 probabilities + row sums?
// Variants of the exp_pack step of csrc/attention_bf16.cu, timed with clock64() for 1 and 2 warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probes/exp_probe tools/probes/exp_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)), "l"(*reinterpret_cast<uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }

// A: the shipped form (pair by pair)
__device__ __forceinline__ void exp_pack_A(const uint32_t (&v)[32], float sc, float nm, float2& la, float2& lb, uint32_t (&pk)[16]) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(nm, nm);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = ffma2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
    const float2 e = make_float2(ex2a(x.x), ex2a(x.y));
    pk[i] = pack_bf16(e.x, e.y);
    if (i & 1) lb = fadd2(lb, e); else la = fadd2(la, e);
  }
}
// B: all exponentials first, consumers afterwards
__device__ __forceinline__ void exp_pack_B(const uint32_t (&v)[32], float sc, float nm, float2& la, float2& lb, uint32_t (&pk)[16]) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(nm, nm);
  float2 e[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = ffma2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
    e[i] = make_float2(ex2v(x.x), ex2v(x.y));
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    pk[i] = pack_bf16(e[i].x, e[i].y);
    if (i & 1) lb = fadd2(lb, e[i]); else la = fadd2(la, e[i]);
  }
}
// C: consumers of pair i placed behind the exponentials of pair i + D (software pipeline written out), volatile asm
template <int D>
__device__ __forceinline__ void exp_pack_C(const uint32_t (&v)[32], float sc, float nm, float2& la, float2& lb, uint32_t (&pk)[16]) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(nm, nm);
  float2 e[16];
#pragma unroll
  for (int i = 0; i < 16 + D; ++i) {
    if (i < 16) {
      const float2 x = ffma2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
      e[i] = make_float2(ex2v(x.x), ex2v(x.y));
    }
    if (i >= D) {
      const int k = i - D;
      uint32_t r;
      // cvt as volatile asm: stays behind the exponentials issued before it
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e[k].y), "f"(e[k].x));
      pk[k] = r;
      if (k & 1) lb = fadd2(lb, e[k]); else la = fadd2(la, e[k]);
    }
  }
}
// G<D>: the consumers of pair k are made to DEPEND on the exponentials of pair k + D (an FFMA2 that adds 0 * e[k+D]:
// not foldable, exact for finite e), so that ptxas -- which otherwise schedules every consumer one pair behind its
// producers whatever the source order -- has to keep D pairs of MUFU results in flight
template <int D>
__device__ __forceinline__ void exp_pack_G(const uint32_t (&v)[32], float sc, float nm, float2& la, float2& lb, uint32_t (&pk)[16]) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(nm, nm), zero2 = make_float2(0.f, 0.f);
  float2 e[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = ffma2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
    e[i] = make_float2(ex2a(x.x), ex2a(x.y));
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float2 ek = (k + D < 16) ? ffma2(e[k + D], zero2, e[k]) : e[k];
    pk[k] = pack_bf16(ek.x, ek.y);
    if (k & 1) lb = fadd2(lb, ek); else la = fadd2(la, ek);
  }
}
// F: like A but without the row sums (the sum would come from a ones column in the PV MMA)
__device__ __forceinline__ void exp_pack_F(const uint32_t (&v)[32], float sc, float nm, float2& la, float2& lb, uint32_t (&pk)[16]) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(nm, nm);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = ffma2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
    pk[i] = pack_bf16(ex2a(x.x), ex2a(x.y));
  }
}


// H: both 32-score chunks of a step in ONE loop (two independent chains side by side): ptxas keeps every consumer one
// ITERATION behind its producers, which is now four MUFUs (32 pipe cycles), not two
__device__ __forceinline__ void exp_pack_H(const uint32_t (&va)[32], const uint32_t (&vb)[32], float sc, float nm, float2& la, float2& lb,
                                           uint32_t (&pa)[16], uint32_t (&pb)[16]) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(nm, nm);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 xa = ffma2(make_float2(__uint_as_float(va[2 * i]), __uint_as_float(va[2 * i + 1])), sc2, nm2);
    const float2 xb = ffma2(make_float2(__uint_as_float(vb[2 * i]), __uint_as_float(vb[2 * i + 1])), sc2, nm2);
    const float2 ea = make_float2(ex2a(xa.x), ex2a(xa.y));
    const float2 eb = make_float2(ex2a(xb.x), ex2a(xb.y));
    pa[i] = pack_bf16(ea.x, ea.y);
    pb[i] = pack_bf16(eb.x, eb.y);
    la = fadd2(la, ea);
    lb = fadd2(lb, eb);
  }
}
// I: as H with four chains (quarters of 16 scores)
__device__ __forceinline__ void exp_pack_I(const uint32_t (&va)[32], const uint32_t (&vb)[32], float sc, float nm, float2& la, float2& lb,
                                           uint32_t (&pa)[16], uint32_t (&pb)[16]) {
  const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(nm, nm);
  float2 lc = make_float2(0.f, 0.f), ld = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 x0 = ffma2(make_float2(__uint_as_float(va[2 * i]), __uint_as_float(va[2 * i + 1])), sc2, nm2);
    const float2 x1 = ffma2(make_float2(__uint_as_float(va[16 + 2 * i]), __uint_as_float(va[16 + 2 * i + 1])), sc2, nm2);
    const float2 x2 = ffma2(make_float2(__uint_as_float(vb[2 * i]), __uint_as_float(vb[2 * i + 1])), sc2, nm2);
    const float2 x3 = ffma2(make_float2(__uint_as_float(vb[16 + 2 * i]), __uint_as_float(vb[16 + 2 * i + 1])), sc2, nm2);
    const float2 e0 = make_float2(ex2a(x0.x), ex2a(x0.y));
    const float2 e1 = make_float2(ex2a(x1.x), ex2a(x1.y));
    const float2 e2 = make_float2(ex2a(x2.x), ex2a(x2.y));
    const float2 e3 = make_float2(ex2a(x3.x), ex2a(x3.y));
    pa[i] = pack_bf16(e0.x, e0.y);
    pa[8 + i] = pack_bf16(e1.x, e1.y);
    pb[i] = pack_bf16(e2.x, e2.y);
    pb[8 + i] = pack_bf16(e3.x, e3.y);
    la = fadd2(la, e0); lb = fadd2(lb, e1); lc = fadd2(lc, e2); ld = fadd2(ld, e3);
  }
  la = fadd2(la, lc); lb = fadd2(lb, ld);
}

template <int VAR>
__global__ void __launch_bounds__(512, 1) probe(const float* in, uint32_t* out, long long* cycles, int iters) {
  uint32_t va[32], vb[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { va[i] = __float_as_uint(in[(threadIdx.x * 64 + i) & 4095]); vb[i] = __float_as_uint(in[(threadIdx.x * 64 + 32 + i) & 4095]); }
  float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
  uint32_t acc = 0;
  float nm = -1.0f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[16];
#define RUN(V)                                                              \
    if (VAR == 0) exp_pack_A(V, 0.125f, nm, la, lb, pk);                     \
    else if (VAR == 1) exp_pack_B(V, 0.125f, nm, la, lb, pk);                \
    else if (VAR == 2) exp_pack_C<2>(V, 0.125f, nm, la, lb, pk);             \
    else if (VAR == 3) exp_pack_C<3>(V, 0.125f, nm, la, lb, pk);             \
    else if (VAR == 4) exp_pack_C<4>(V, 0.125f, nm, la, lb, pk);             \
    else if (VAR == 5) exp_pack_G<2>(V, 0.125f, nm, la, lb, pk);                \
    else if (VAR == 7) exp_pack_G<3>(V, 0.125f, nm, la, lb, pk);                \
    else if (VAR == 8) exp_pack_G<4>(V, 0.125f, nm, la, lb, pk);                \
    else exp_pack_F(V, 0.125f, nm, la, lb, pk);
    if (VAR == 9 || VAR == 10) {
      uint32_t pk2[16];
      if (VAR == 9) exp_pack_H(va, vb, 0.125f, nm, la, lb, pk, pk2);
      else exp_pack_I(va, vb, 0.125f, nm, la, lb, pk, pk2);
#pragma unroll
      for (int i = 0; i < 16; ++i) { acc ^= pk[i]; acc += pk2[i]; }
    } else {
    RUN(va)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= pk[i];
    RUN(vb)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += pk[i];
    }
    nm -= 1e-3f;
    // keep the inputs "live and changing" so nothing is hoisted out of the loop
    va[it & 31] ^= acc & 1u;
    vb[(it + 7) & 31] ^= (acc >> 1) & 1u;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(la.x + la.y + lb.x + lb.y);
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int VAR>
void run(const char* name, const float* in, uint32_t* out, long long* cyc) {
  const int iters = 2000;
  for (int threads : {128, 256, 512}) {
    probe<VAR><<<148, threads>>>(in, out, cyc, iters);
    cudaDeviceSynchronize();
    probe<VAR><<<148, threads>>>(in, out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%-44s warps/SMSP %d: %7.1f cycles per 64 scores per warp  (%s)\n", name, threads / 128, double(c) / iters, cudaGetErrorString(e));
  }
}

int main() {
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = float((i * 37) % 101) * 0.05f - 3.0f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("A pair by pair (shipped)", in, out, cyc);
  run<1>("B all ex2 first, consumers after", in, out, cyc);
  run<2>("C consumers 2 pairs behind (volatile)", in, out, cyc);
  run<3>("C consumers 3 pairs behind (volatile)", in, out, cyc);
  run<4>("C consumers 4 pairs behind (volatile)", in, out, cyc);
  run<5>("G consumers depend on pair k+2", in, out, cyc);
  run<7>("G consumers depend on pair k+3", in, out, cyc);
  run<8>("G consumers depend on pair k+4", in, out, cyc);
  run<6>("F no row sums", in, out, cyc);
  run<9>("H two chunks in one loop", in, out, cyc);
  run<10>("I four quarter-chunks in one loop", in, out, cyc);
  return 0;
}
