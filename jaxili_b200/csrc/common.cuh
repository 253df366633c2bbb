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

// programmatic dependent launch (PDL): wait for the preceding grid in the stream (no-op without the launch attribute);
// allow the following grid to start launching
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
#ifndef MAFB_FFMA2
#define MAFB_FFMA2 1
#endif
// (d0, d1) += (a0, a1) * (b0, b1): sm_100 packed FP32 FMA (SASS FFMA2).  ptxas keeps the operands in aligned
// register pairs (they come from 128-bit shared-memory loads) and folds b0 == b1 into a scalar-broadcast operand.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// in: [Q][RP] feature-major, m: [Q][CP] row-major.  Per q: one TM-wide and one TN-wide vector
// load feed TM*TN FFMAs.  Only q = q0, q0+qs, q0+2qs, ... are visited (lane-group split of the
// reduction; qs = 1 for the whole sum).
template <int TM, int TN>
__device__ __forceinline__ void mm_outer_acc(float (&acc)[TN][TM], const float* __restrict__ in, int Q,
                                             int RP, const float* __restrict__ m, int CP, int r0, int c0,
                                             int q0, int qs) {
#pragma unroll
  for (int j = 0; j < TN; ++j)
#pragma unroll
    for (int i = 0; i < TM; ++i) acc[j][i] = 0.f;
  const float* a = in + r0 + q0 * RP;
  const float* w = m + c0 + q0 * CP;
  const int da = qs * RP, dw = qs * CP;
#pragma unroll 4
  for (int q = q0; q < Q; q += qs) {
    float av[TM], wv[TN];
    ldv(av, a);
    ldv(wv, w);
    a += da;
    w += dw;
#if MAFB_FFMA2
    // packed FFMA2: two accumulators along the row pair per instruction, the weight as a scalar broadcast
    // operand -- same FP32 FMA rate as FFMA with half the issue slots (tests/micro/ffma2_peak.cu), bit-identical
    static_assert(TM % 2 == 0, "row tile must be even for the packed form");
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int i = 0; i < TM; i += 2) ffma2(acc[j][i], acc[j][i + 1], av[i], av[i + 1], wv[j], wv[j]);
#else
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int i = 0; i < TM; ++i) acc[j][i] = fmaf(av[i], wv[j], acc[j][i]);
#endif
  }
}

template <int TM, int TN, class Epi>
__device__ __forceinline__ void mm_outer(const float* __restrict__ in, int Q, int RP,
                                         const float* __restrict__ m, int CP, int r0, int c0,
                                         Epi&& epi) {
  float acc[TN][TM];
  mm_outer_acc<TM, TN>(acc, in, Q, RP, m, CP, r0, c0, 0, 1);
  epi(acc);
}

// 4-way lane-group split of the reduction for a 4x4 tile.  The four lanes {l, l^8, l^16, l^24}
// hold partial sums of the same tile (kq = (lane>>3)&3 selects q = kq, kq+4, ...).  A warp-shuffle
// reduce-scatter leaves lane kq with the complete sums of tile column kq (4 rows) in res[].
__device__ __forceinline__ void reduce_scatter4(const float (&acc)[4][4], int kq, float (&res)[4]) {
  const bool b1 = (kq & 2) != 0, b0 = (kq & 1) != 0;
  float t[2][4];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = b1 ? acc[c][i] : acc[2 + c][i];
      const float mine = b1 ? acc[2 + c][i] : acc[c][i];
      t[c][i] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b0 ? t[0][i] : t[1][i];
    const float mine = b0 ? t[1][i] : t[0][i];
    res[i] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
  }
}

// 2-way variant: lanes {l, l^16} hold partial sums of the same 4x4 tile (kq = lane>>4 takes
// q = kq, kq+2, ...); lane kq ends with complete sums of tile columns 2kq, 2kq+1 in res[2][4].
__device__ __forceinline__ void reduce_scatter2(const float (&acc)[4][4], int kq, float (&res)[2][4]) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = kq ? acc[c][i] : acc[2 + c][i];
      const float mine = kq ? acc[2 + c][i] : acc[c][i];
      res[c][i] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
    }
}

// dW[k][n] = sum_r A[row(k)][r] * G[n][r] for the 4x4 strided tile k = kb + i*nKB, n = nb + j*nNB,
// k in [0, K]: row(k) = k for k < K and ones_row for k == K (a feature row of ones, which makes
// dW[K][n] the bias gradient; the flat layout stores the bias right behind the kernel).
// The row dimension is split over the two half-warps (rh = lane>>4 takes r = 4rh, 4rh+8, ...) and
// combined with a warp-shuffle reduce-scatter: half rh ends up owning tile rows i = 2rh, 2rh+1.
// The strided k/n assignment makes the row loads of one warp instruction hit consecutive feature
// rows, which the RP pitch maps to distinct banks.
__device__ __forceinline__ void mm_inner_rs2(const float* __restrict__ A, int K, int ones_row,
                                             const float* __restrict__ G, int N, int Rt, int RP, int kb,
                                             int nKB, int nb, int nNB, int rh, bool store,
                                             float* __restrict__ out, bool accumulate) {
  float acc[4][4];
#if MAFB_FFMA2
  float acc_odd[4][4];
#endif
  const float* ap[4];
  const float* gp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = kb + i * nKB;
    ap[i] = A + (k < K ? k : ones_row) * RP + 4 * rh;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[i][j] = 0.f;
#if MAFB_FFMA2
      acc_odd[i][j] = 0.f;
#endif
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = nb + j * nNB;
    gp[j] = G + (n < N ? n : N - 1) * RP + 4 * rh;
  }
#pragma unroll 2
  for (int r = 4 * rh; r < Rt; r += 8) {
    float av[4][4], gv[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ldv(av[i], ap[i]); ap[i] += 8; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { ldv(gv[j], gp[j]); gp[j] += 8; }
#if MAFB_FFMA2
    // packed FFMA2 over the row pairs (t, t+1): even rows accumulate in acc, odd rows in acc_odd
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int t = 0; t < 4; t += 2)
          ffma2(acc[i][j], acc_odd[i][j], av[i][t], av[i][t + 1], gv[j][t], gv[j][t + 1]);
#else
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[i][j] = fmaf(av[i][t], gv[j][t], acc[i][j]);
#endif
  }
#if MAFB_FFMA2
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] += acc_odd[i][j];
#endif
  // reduce-scatter across the half-warps: keep tile rows 2rh, 2rh+1
  float res[2][4];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float send = rh ? acc[c][j] : acc[2 + c][j];
      const float mine = rh ? acc[2 + c][j] : acc[c][j];
      res[c][j] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  if (store) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int k = kb + (2 * rh + c) * nKB;
      if (k <= K) {
        float* prow = out + k * N;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = nb + j * nNB;
          if (n < N) {
            // this thread is the only writer of its entries of the CTA's partial-gradient slice;
            // later tiles add with a fire-and-forget reduction (no load latency, same order each run)
            if (accumulate) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(prow + n), "f"(res[c][j]) : "memory");
            else prow[n] = res[c][j];
          }
        }
      }
    }
  }
}

}  // namespace mafb
