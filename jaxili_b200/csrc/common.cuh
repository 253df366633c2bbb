// common.cuh -- device helpers shared by the fused MAF kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "plan.h"

namespace mafb {

// ---------------------------------------------------------------- mbarrier + bulk copy (TMA 1-D)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (SASS: UBLKCP), completion counted in bytes on `bar`.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// order prior generic-proxy accesses to shared memory before subsequent async-proxy writes
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- small vector loads / stores
__device__ __forceinline__ void ldv(float (&d)[4], const float* p) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}
__device__ __forceinline__ void ldv(float (&d)[2], const float* p) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  d[0] = v.x; d[1] = v.y;
}
__device__ __forceinline__ void ldv(float (&d)[1], const float* p) { d[0] = *p; }
__device__ __forceinline__ void stv(float* p, const float (&d)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
}
__device__ __forceinline__ void stv(float* p, const float (&d)[2]) {
  *reinterpret_cast<float2*>(p) = make_float2(d[0], d[1]);
}
__device__ __forceinline__ void stv(float* p, const float (&d)[1]) { *p = d[0]; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- activations (jax.nn semantics)
// returns h = act(a) and d = act'(a)
template <bool RELU>
__device__ __forceinline__ void act_eval(int act, float a, float& h, float& d) {
  if (RELU) {
    h = fmaxf(a, 0.f);
    d = a > 0.f ? 1.f : 0.f;
    return;
  }
  switch (act) {
    case 0: h = fmaxf(a, 0.f); d = a > 0.f ? 1.f : 0.f; break;
    case 1: {  // silu
      const float s = 1.f / (1.f + expf(-a));
      h = a * s;
      d = s * (1.f + a * (1.f - s));
    } break;
    case 2: {  // tanh
      const float t = tanhf(a);
      h = t;
      d = 1.f - t * t;
    } break;
    case 3: {  // gelu, tanh approximation
      const float c = 0.7978845608028654f;
      const float in = c * (a + 0.044715f * a * a * a);
      const float t = tanhf(in);
      h = 0.5f * a * (1.f + t);
      d = 0.5f * (1.f + t) + 0.5f * a * (1.f - t * t) * c * (1.f + 3.f * 0.044715f * a * a);
    } break;
    case 4: {  // elu, alpha = 1
      const float e = expm1f(fminf(a, 0.f));
      h = a > 0.f ? a : e;
      d = a > 0.f ? 1.f : e + 1.f;
    } break;
    case 5: {  // sigmoid
      const float s = 1.f / (1.f + expf(-a));
      h = s;
      d = s * (1.f - s);
    } break;
    case 6: {  // softplus = logaddexp(a, 0)
      h = fmaxf(a, 0.f) + log1pf(expf(-fabsf(a)));
      d = 1.f / (1.f + expf(-a));
    } break;
    default: {  // leaky_relu 0.01
      h = a >= 0.f ? a : 0.01f * a;
      d = a >= 0.f ? 1.f : 0.01f;
    } break;
  }
}

// ---------------------------------------------------------------- register-tiled micro-GEMMs on
// shared-memory operands.  Activations and gradients are FEATURE-MAJOR: buf[feature][row] with
// row pitch RP (RP/4 odd, so consecutive feature rows start 16 bytes apart modulo 128).

// out[c][r] = sum_q in[q][r] * m[q][c]   for r in r0..r0+TM-1, c in c0..c0+TN-1
// in: [Q][RP] feature-major, m: [Q][CP] row-major.  Per q: one TM-wide and one TN-wide vector
// load feed TM*TN FFMAs.  Used for the forward linears (m = masked W) and for the
// input-gradient products (in = G, m = masked W^T).
template <int TM, int TN, class Epi>
__device__ __forceinline__ void mm_outer(const float* __restrict__ in, int Q, int RP,
                                         const float* __restrict__ m, int CP, int r0, int c0,
                                         Epi&& epi) {
  float acc[TN][TM];
#pragma unroll
  for (int j = 0; j < TN; ++j)
#pragma unroll
    for (int i = 0; i < TM; ++i) acc[j][i] = 0.f;
  const float* a = in + r0;
  const float* w = m + c0;
#pragma unroll 4
  for (int q = 0; q < Q; ++q) {
    float av[TM], wv[TN];
    ldv(av, a);
    ldv(wv, w);
    a += RP;
    w += CP;
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int i = 0; i < TM; ++i) acc[j][i] = fmaf(av[i], wv[j], acc[j][i]);
  }
  epi(acc);
}

// dW[k][n] = sum_r A[k][r] * G[n][r]  for the TK x TN strided tile k = kb + i*nKB, n = nb + j*nNB.
// The strided assignment makes the TK (TN) row loads of one warp instruction hit consecutive
// feature rows, which the RP pitch maps to distinct banks.
template <int TK, int TN>
__device__ __forceinline__ void mm_inner(const float* __restrict__ A, int K, const float* __restrict__ G,
                                         int N, int Rt, int RP, int kb, int nKB, int nb, int nNB,
                                         float* __restrict__ out, bool accumulate) {
  float acc[TK][TN];
  const float* ap[TK];
  const float* gp[TN];
#pragma unroll
  for (int i = 0; i < TK; ++i) {
    const int k = kb + i * nKB;
    ap[i] = A + (k < K ? k : K - 1) * RP;
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = nb + j * nNB;
    gp[j] = G + (n < N ? n : N - 1) * RP;
  }
#pragma unroll 2
  for (int r = 0; r < Rt; r += 4) {
    float av[TK][4], gv[TN][4];
#pragma unroll
    for (int i = 0; i < TK; ++i) ldv(av[i], ap[i] + r);
#pragma unroll
    for (int j = 0; j < TN; ++j) ldv(gv[j], gp[j] + r);
#pragma unroll
    for (int i = 0; i < TK; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[i][j] = fmaf(av[i][t], gv[j][t], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < TK; ++i) {
    const int k = kb + i * nKB;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = nb + j * nNB;
      if (k < K && n < N) {
        float* p = out + k * N + n;
        *p = accumulate ? *p + acc[i][j] : acc[i][j];
      }
    }
  }
}

}  // namespace mafb
