// chain_kernel.cuh -- warp-specialised fused ConditionalMAF kernel for sm_100a (the default flow kernel).
//
// Same contract, shared-memory plan and weight images as flow_kernel.cuh (forward: jaxili/model.py:1092-1098,
// 903-921, 848-874, 718-743, 665-686, 517-544; reverse: what jax.value_and_grad produces for loss.py:35-83),
// different execution structure.  At the batch sizes of the narrow configs an SM sees ~28 rows, and a MADE is a
// chain of 15 forward + 15 reverse dependent linears: a design that spreads every linear over the whole CTA pays
// a CTA barrier, a pipeline fill and a drain per linear (47 % barrier stall in round 1).  Here
//
//   * CHAIN warps own rows.  A PAIR of warps (c, c+4: same SM sub-partition) holds rows 8p..8p+7 through every
//     linear of every layer, forward and reverse; the two warps split the output columns of a product and meet
//     at a 64-thread named barrier -- the only synchronisation on the dependent chain.  Two chain warps per
//     sub-partition hide each other's latencies (tests/micro/lone_warp.cu: one warp alone runs the inner loop at
//     51 cycles per 16 FFMA2, two reach the pipe's 37.5); warp ids: producer, chain warps, gradient warps.
//     A lane holds 8 rows x 4 columns and one of 4 parts of the reduction (32 accumulators, 3 loads per 16 packed
//     FFMA2, operands prefetched through a register ring); a two-stage shuffle reduce-scatter leaves every lane
//     2 columns x 2 rows per row block for the epilogue.
//   * GRADIENT warps compute the weight/bias gradients [A;1]^T G, the only product that crosses rows: once per
//     layer, all linears of the layer as one flat list of 8x4 tiles (one tile per lane, reduction over all rows
//     of the tile, no cross-lane step).  They consume the gradient buffers asynchronously (two buffer sets,
//     named barriers full[b] / empty[b]) and fill the issue slots the chain leaves empty.
//   * one PRODUCER warp streams the weight images and the input tiles with bulk copies (TMA 1-D) behind
//     full/empty mbarriers.
#pragma once
#include "common.cuh"
#include "flow_kernel.cuh"

namespace mafb {

constexpr int CH_CW = 8;                         // chain warps: 4 pairs, 8 rows per pair, 32-row tiles
constexpr int CH_GW = 6;                         // gradient warps (one 8x4 tile per lane covers a [50,50] layer in one round)
constexpr int CH_NT_BWD = 32 * (1 + CH_GW + CH_CW);
constexpr int CH_NT_FWD = 32 * (1 + CH_CW);
constexpr int CH_BAR_FULL = 1;                   // named barriers: full[b] = 1 + b, empty[b] = 3 + b
constexpr int CH_BAR_EMPTY = 3;
constexpr int CH_BAR_TILE = 5;                   // gradient warps are done with the tile's stash
constexpr int CH_BAR_CHAIN = 6;                  // chain warps only (loss hand-off at the end)
constexpr int CH_BAR_PAIR = 7;                   // + pair index: the two warps of a row pair

__device__ __forceinline__ void nbar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void nbar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// producer-side wait: poll with a back-off so the spin does not eat issue slots of the sub-partition it shares
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(256);
  }
}

// acc[j][i] += a[i] * w[j]: 8 packed FFMA2 (row pairs, the weight as the broadcast operand)
__device__ __forceinline__ void ch_fma_tile(float (&acc)[4][4], const float (&a)[4], const float (&w)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; i += 2) ffma2(acc[j][i], acc[j][i + 1], a[i], a[i + 1], w[j], w[j]);
}

// n reduction steps of the lane's 8x4 tile (two 4x4 tiles sharing the weight operand); pa / pw advance by da / dw
// floats per step.  An odd step is peeled off the front; the rest runs through a 2-stage register ring: the loads
// of step s+2 are issued behind the FMAs of step s, one step (16 FFMA2 of this warp plus the partner warp's
// share of the pipe) ahead of their use.
__device__ __forceinline__ void ch_mac(float (&acc)[2][4][4], const float* __restrict__ pa,
                                       const float* __restrict__ pw, int n, int da, int dw) {
  if (n & 1) {
    float a0[4], a1[4], w[4];
    ldv(a0, pa);
    ldv(a1, pa + 4);
    ldv(w, pw);
    ch_fma_tile(acc[0], a0, w);
    ch_fma_tile(acc[1], a1, w);
    pa += da;
    pw += dw;
  }
  int m = n & ~1;
  if (m == 0) return;
  float a0[2][4], a1[2][4], w[2][4];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    ldv(a0[u], pa + u * da);
    ldv(a1[u], pa + u * da + 4);
    ldv(w[u], pw + u * dw);
  }
#pragma unroll 1
  for (m -= 2; m > 0; m -= 2) {
    pa += 2 * da;
    pw += 2 * dw;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      ch_fma_tile(acc[0], a0[u], w[u]);
      ch_fma_tile(acc[1], a1[u], w[u]);
      ldv(a0[u], pa + u * da);
      ldv(a1[u], pa + u * da + 4);
      ldv(w[u], pw + u * dw);
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    ch_fma_tile(acc[0], a0[u], w[u]);
    ch_fma_tile(acc[1], a1[u], w[u]);
  }
}

// Reduce-scatter stages of a 4x4 tile (acc[col][row]) over the 4 lanes that hold its reduction parts:
// stage 1 (partner lane ^ 16) keeps column pair `hi`, stage 2 (lane ^ 8) row pair `hi`.
__device__ __forceinline__ void ch_rs_cols(const float (&acc)[4][4], bool hi, float (&t)[2][4]) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = hi ? acc[c][i] : acc[2 + c][i];
      const float mine = hi ? acc[2 + c][i] : acc[c][i];
      t[c][i] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
    }
}
__device__ __forceinline__ void ch_rs_rows(const float (&t)[2][4], bool hi, float (&r)[2][2]) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = hi ? t[c][i] : t[c][2 + i];
      const float mine = hi ? t[c][2 + i] : t[c][i];
      r[c][i] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
    }
}

// One warp's share of a product of the chain, for the pair's 8 rows: out[c][r] = sum_q in[q][r] * W[q][c] for the
// column blocks first .. first + nblk - 1.  lane = kq * 8 + column slot; the lane's reduction part is
// q = kq, kq + 4, ...  After the reduce-scatter the lane owns columns (ce, ce+1) x 2 rows per row block:
// epi(ce, rr, v) with v[2 cols][2 rows], rr = first row relative to the pair's 8 rows.
template <class Epi>
__device__ __forceinline__ void chain_gemm(const float* __restrict__ in, int RP, const float* __restrict__ W, int ldw,
                                           int first, int nblk, int nsteps, int lane, Epi&& epi) {
  const int cbl = lane & 7;
  const int kq = lane >> 3;
  const bool b1 = (lane & 16) != 0, b0 = (lane & 8) != 0;
#pragma unroll 1
  for (int base = 0; base < nblk; base += 8) {
    float acc[2][4][4];
    const int cb = base + cbl;
    const bool ok = cb < nblk;
    const int c0 = 4 * (first + (ok ? cb : nblk - 1));
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[g][j][i] = 0.f;
    ch_mac(acc, in + kq * RP, W + kq * ldw + c0, nsteps, 4 * RP, 4 * ldw);
    const int ce = c0 + (b1 ? 2 : 0);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float t[2][4], r[2][2];
      ch_rs_cols(acc[g], b1, t);
      ch_rs_rows(t, b0, r);
      if (ok) epi(ce, 4 * g + (b0 ? 2 : 0), r);
    }
  }
}

// Weight + bias gradient tile of one lane: out[i][j] = sum_r A[i * a_stride][r] * G[j * g_stride][r] over the
// tile's rows, i < 8, j < 4.  Packed FFMA2 over the row pairs (t, t+1): even rows accumulate in acc, odd rows in
// acc_odd.  A's feature rows continue into the ones row right behind the last feature, which makes the last
// row of the product the bias gradient.
__device__ __forceinline__ void dw_tile(const float* __restrict__ A, int a_stride, const float* __restrict__ G,
                                        int g_stride, int Rt, float (&out)[8][4]) {
  float acc_odd[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { out[i][j] = 0.f; acc_odd[i][j] = 0.f; }
#pragma unroll 1
  for (int r = 0; r < Rt; r += 4) {
    float av[8][4], gv[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ldv(gv[j], G + j * g_stride + r);
#pragma unroll
    for (int i = 0; i < 8; ++i) ldv(av[i], A + i * a_stride + r);
    // row pair outermost: 32 independent FFMA2 between two updates of the same accumulator pair
#pragma unroll
    for (int t = 0; t < 4; t += 2)
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ffma2(out[i][j], acc_odd[i][j], av[i][t], av[i][t + 1], gv[j][t], gv[j][t + 1]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[i][j] += acc_odd[i][j];
}

// NLIN > 0 fixes the number of masked linears per MADE at compile time (the loops over them unroll and the
// per-linear descriptors become constant-bank operands); NLIN == 0 reads it from the plan.
template <bool BWD, bool RELU, int NLIN>
__global__ void __launch_bounds__(BWD ? CH_NT_BWD : CH_NT_FWD, 1)
maf_chain_kernel(const __grid_constant__ Plan P, const __grid_constant__ FlowArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
#ifdef MAFB_CTA_TIMES
  if (A.trace && threadIdx.x == 0) {   // debug: %globaltimer at entry of every CTA
    unsigned long long tns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
    A.trace[2 * blockIdx.x] = (long long)tns;
  }
#endif
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem_raw);   // [2] weight image landed
  uint64_t* wempty = wfull + 2;                              // [2] all chain warps are done with it
  uint64_t* sfull = wfull + 4;                               // [2] input tile landed
  uint64_t* sempty = wfull + 6;                              // [2] all chain warps have transposed it
  float* loss_red = reinterpret_cast<float*>(smem_raw + 64); // [CH_CW]
  float* sm = reinterpret_cast<float*>(smem_raw + MAFB_SMEM_HDR);

  constexpr int NT = BWD ? CH_NT_BWD : CH_NT_FWD;
  constexpr int GW = BWD ? CH_GW : 0;
  constexpr int NB = 32 * (GW + CH_CW);                      // threads on the full / empty / tile barriers
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifndef MAFB_EXP_GRADFIRST
  // warp 0 = producer, then the chain warps, then the gradient warps: with the gradient warps on the highest warp
  // ids the training pass is 2.1 us (4.5 %) shorter than with the chain warps there (c2; tests/cta_times.py)
  constexpr int W_CHAIN0 = 1, W_GRAD0 = 1 + CH_CW;
#else
  constexpr int W_GRAD0 = 1, W_CHAIN0 = 1 + GW;
#endif
  const int D = P.D, C = P.C, L = P.L, RP = P.RP;
  constexpr int UL = NLIN > 0 ? NLIN : 1;                    // unroll factor of the loops over the linears
  const int n_lin = NLIN > 0 ? NLIN : P.n_lin;
  const int Rt = A.Rt;
  const int grid = gridDim.x;
  const int my_n = (A.n_tiles - (int)blockIdx.x + grid - 1) / grid;
  const int per_tile = BWD ? 2 * L : L;
  const int total_req = my_n * per_tile;
  const bool two_wbuf = P.n_wbuf == 2;

#ifdef MAFB_TRACE
  // CTA 0: chain warp 0 stamps into trace[0..499], gradient warp 0 into trace[500..999] ((id, clock) pairs)
  int n_trace = 0;
  const bool tr_on = A.trace && blockIdx.x == 0 && lane == 0 && (warp == W_CHAIN0 || (BWD && warp == W_GRAD0));
  long long* const tbase = A.trace + ((BWD && warp == W_GRAD0) ? 500 : 0);
  auto stamp = [&](int id) {
    if (tr_on && n_trace < 248) {
      tbase[2 * n_trace] = id;
      tbase[2 * n_trace + 1] = clock64();
      ++n_trace;
    }
  };
  auto stamp_end = [&]() { if (tr_on) tbase[498] = n_trace; };
#else
  auto stamp = [](int) {};
  auto stamp_end = []() {};
#endif

  if (tid == 0) {
    mbar_init(&wfull[0], 1);
    mbar_init(&wfull[1], 1);
    mbar_init(&wempty[0], CH_CW);
    mbar_init(&wempty[1], CH_CW);
    mbar_init(&sfull[0], 1);
    mbar_init(&sfull[1], 1);
    mbar_init(&sempty[0], CH_CW);
    mbar_init(&sempty[1], CH_CW);
    mbar_fence_init();
  }
  // activations, gradients and their pad rows start at zero (pad rows meet zero weights, and 0 * garbage may be NaN)
  {
    const int n4 = (P.smem_bytes - MAFB_SMEM_HDR) / 16;
    float4* z = reinterpret_cast<float4*>(sm);
    for (int idx = P.o_xin / 4 + tid; idx < n4; idx += NT) z[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // everything above is independent of the preceding kernel in the stream (the optimiser step that rewrites the
  // weight images and bumps the step counter): under programmatic dependent launch it overlaps that kernel's tail
  griddep_wait();
  const long long row_base = A.batch_index ? (long long)(A.batch_index[0] % A.n_batches) * A.B : 0;
  __syncthreads();
  float* const cst = sm + P.o_cst;   // [shift D][inv_scale D][mean C][std C]
  for (int idx = tid; idx < 2 * D + 2 * C; idx += NT) cst[idx] = A.stdc[idx];
  if (BWD) {
    // a feature row of ones right behind the last feature of every weight-gradient A operand (row K; it meets a
    // zero row of the weight image in the forward products): [A;1]^T G yields the bias gradient as row K
    for (int idx = tid; idx < RP; idx += NT) {
      for (int l = 0; l < L; ++l) {
        sm[P.o_xin + l * P.xin_stride + P.lin[0].K * RP + idx] = 1.f;
        for (int i = 1; i < n_lin; ++i)
          sm[P.o_slot + l * P.slot_stride + P.h_off[i] + P.lin[i].K * RP + idx] = 1.f;
      }
    }
  }
  __syncthreads();
  stamp(0);

  // ======================================================================== producer warp
  if (warp == 0) {
    auto stage_issue = [&](int i) {
      const int t = blockIdx.x + i * grid;
      const long long row0 = (long long)t * Rt;
      const int valid = (int)min((long long)Rt, A.B - row0);
      const float* xs = A.x + (row_base + row0) * D;
      const float* ys = A.y + (row_base + row0) * C;
      float* dst = sm + P.o_stage + (i & 1) * P.stage_stride;
      const bool fast = (valid % 4 == 0) && ((reinterpret_cast<uintptr_t>(xs) & 15) == 0) &&
                        (C == 0 || (reinterpret_cast<uintptr_t>(ys) & 15) == 0);
      if (fast) {
        if (lane == 0) {
          mbar_expect_tx(&sfull[i & 1], (uint32_t)(valid * (D + C) * 4));
          bulk_g2s(dst, xs, (uint32_t)(valid * D * 4), &sfull[i & 1]);
          if (C) bulk_g2s(dst + P.R * D, ys, (uint32_t)(valid * C * 4), &sfull[i & 1]);
        }
      } else {
        for (int idx = lane; idx < valid * D; idx += 32) dst[idx] = xs[idx];
        for (int idx = lane; idx < valid * C; idx += 32) dst[P.R * D + idx] = ys[idx];
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfull[i & 1]);
      }
    };
    if (my_n > 0) stage_issue(0);
    if (my_n > 1) stage_issue(1);
    const int jstage = per_tile > 2 ? 2 : per_tile - 1;
    int ti = 0, j = 0;
    for (int s = 0; s < total_req; ++s) {
      const int buf = two_wbuf ? (s & 1) : 0;
      const int use = two_wbuf ? (s >> 1) : s;         // earlier fills of this buffer
      if (lane == 0) {
        if (use > 0) mbar_wait_sleep(&wempty[buf], (uint32_t)((use - 1) & 1));
        int layer, part;
        if (!BWD || j < L) { layer = j; part = 0; } else { layer = 2 * L - 1 - j; part = 1; }
        const float* src = A.wimg + (size_t)layer * P.img_floats;
        uint32_t bytes;
        if (part == 0) bytes = P.fwd_floats * 4u;
        else { src += P.fwd_floats; bytes = P.bwd_floats * 4u; }
        mbar_expect_tx(&wfull[buf], bytes);
        bulk_g2s(sm + P.o_wbuf + buf * P.wbuf_stride, src, bytes, &wfull[buf]);
      }
      if (j == jstage && ti + 2 < my_n) {
        // tile ti has been transposed out of its staging buffer (its first weight request was released)
        if (lane == 0) mbar_wait_sleep(&sempty[ti & 1], (uint32_t)((ti >> 1) & 1));
        __syncwarp();
        stage_issue(ti + 2);
      }
      if (++j == per_tile) { j = 0; ++ti; }
    }
    return;
  }

  // ======================================================================== gradient warps
  // layer step sq = ti * L + (L - 1 - k) alternates between the two gradient-buffer sets
  if (BWD && warp >= W_GRAD0 && warp < W_GRAD0 + GW) {
    const int gw = warp - W_GRAD0;
    float* pg_cta = A.partial_grad + (size_t)blockIdx.x * P.pgs;
    const int total_steps = my_n * L;
    // flat list of 8x4 tiles over the linears of one layer, one tile per lane and round; a tile takes the strided
    // features k = kb + i * nKB (i < 8; k == K is the ones row = bias), n = nb + j * nNB (j < 4): the rows one
    // warp instruction reads are consecutive features, which the pitch maps to distinct banks
    int n_tiles_layer = 0;
#pragma unroll UL
    for (int i = 0; i < n_lin; ++i) n_tiles_layer += ((P.lin[i].K + 8) >> 3) * (P.lin[i].NP >> 2);
    int sq = 0;
    for (int ti = 0; ti < my_n; ++ti) {
      const bool accum = ti > 0;
      for (int k = L - 1; k >= 0; --k, ++sq) {
        float* pg = pg_cta + (size_t)k * P.ppl;
        const float* gset = sm + P.o_gb + (sq & 1) * P.gb_stride;
        nbar_sync(CH_BAR_FULL + (sq & 1), NB);
        stamp(500 + 10 * k);
        for (int g0 = gw * 32; g0 < n_tiles_layer; g0 += GW * 32) {
          int g = g0 + lane;
          const bool active = g < n_tiles_layer;
          if (!active) g = 0;
          // which linear, which tile
          int li = 0, nKB = 1, nNB = 1, local = 0, base = 0;
#pragma unroll UL
          for (int i = 0; i < n_lin; ++i) {
            const int kbs = (P.lin[i].K + 8) >> 3, nbs = P.lin[i].NP >> 2;
            if (g >= base && g < base + kbs * nbs) { li = i; nKB = kbs; nNB = nbs; local = g - base; }
            base += kbs * nbs;
          }
          g = local;
          const int kb = g / nNB, nb = g - kb * nNB;
          int K = 0, N = 0, pw_off = 0, goff = 0, aoff = 0;
#pragma unroll UL
          for (int i = 0; i < n_lin; ++i)
            if (i == li) {
              K = P.lin[i].K; N = P.lin[i].N; pw_off = P.lin[i].pw_off; goff = P.g_off[i];
              aoff = i == 0 ? P.o_xin + k * P.xin_stride : P.o_slot + k * P.slot_stride + P.h_off[i];
            }
          float out[8][4];
          dw_tile(sm + aoff + kb * RP, nKB * RP, gset + goff + nb * RP, nNB * RP, Rt, out);
          if (active) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int kk = kb + i * nKB;
              if (kk <= K) {
                float* prow = pg + pw_off + kk * N;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int n = nb + j * nNB;
                  if (n < N) {
                    // this thread is the only writer of its entries of the CTA's partial-gradient slice; later
                    // tiles add with a fire-and-forget reduction (no load latency, same order each run)
                    if (accum) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(prow + n), "f"(out[i][j]) : "memory");
                    else prow[n] = out[i][j];
                  }
                }
              }
            }
          }
        }
        stamp(600 + 10 * k);
        if (sq + 2 < total_steps) nbar_arrive(CH_BAR_EMPTY + (sq & 1), NB);
      }
      if (ti + 1 < my_n) nbar_arrive(CH_BAR_TILE, NB);
    }
    stamp_end();
#ifdef MAFB_CTA_TIMES
    if (A.trace && warp == W_GRAD0 && lane == 0) {
      unsigned long long tns;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
      A.trace[2 * blockIdx.x + 1] = (long long)tns;
    }
#endif
    return;
  }

  // ======================================================================== chain warps
  const int cw = warp - W_CHAIN0;
  const int pr = cw & 3, hf = cw >> 2;                // row pair, column half
  const int nRB = Rt / 4;
  const bool has_rows = 2 * pr < nRB;
  const bool rb1 = 2 * pr + 1 < nRB;                  // the pair's second row block exists in the tile
  const bool own_rb = 2 * pr + hf < nRB;              // this warp's row block for the row-wise steps
  const int r0 = 8 * (has_rows ? pr : 0);
  float* const gu = sm + P.o_gu;
  float* const gx = sm + P.o_gx;
  float* const ld = sm + P.o_ld;
  float my_loss = 0.f;
  int req = 0, sq = 0;

  auto w_acquire = [&](int s) -> const float* {
    const int buf = two_wbuf ? (s & 1) : 0;
    const int use = two_wbuf ? (s >> 1) : s;
    mbar_wait(&wfull[buf], (uint32_t)(use & 1));
    return sm + P.o_wbuf + buf * P.wbuf_stride;
  };
  auto w_release = [&](int s) {
    __syncwarp();
    if (lane == 0) mbar_arrive(&wempty[two_wbuf ? (s & 1) : 0]);
  };
  auto pair_sync = [&]() { nbar_sync(CH_BAR_PAIR + pr, 64); };
  // rows rr, rr+1 of the pair (relative) exist in the tile?  (a row block is all or nothing)
  auto rows_ok = [&](int rr) { return rr < 4 || rb1; };
  // this warp's column blocks of a product: wide products are split between the pair's warps, narrow ones
  // (all blocks fit the 8 column slots of one warp) are done by the first warp alone
  auto my_blocks = [&](const FlowArgs::ChainPh& ph, int& first, int& nblk) {
    if (ph.split) {
      const int na = (ph.nCB + 1) >> 1;
      first = hf ? na : 0;
      nblk = hf ? ph.nCB - na : na;
    } else {
      first = 0;
      nblk = hf ? 0 : ph.nCB;
    }
  };

  for (int ti = 0; ti < my_n; ++ti) {
    const int t = blockIdx.x + ti * grid;
    const long long row0 = (long long)t * Rt;
    const int valid = (int)min((long long)Rt, A.B - row0);

    // the previous tile's activations are still being read by the gradient warps
    if (BWD && ti > 0) nbar_sync(CH_BAR_TILE, NB);

    // ---- stage in: this warp's row block, raw [row][feature] -> standardised feature-major
    stamp(1);
    mbar_wait(&sfull[ti & 1], (uint32_t)((ti >> 1) & 1));
    stamp(2);
    if (own_rb) {
      const float* st = sm + P.o_stage + (ti & 1) * P.stage_stride;
      float* xin0 = sm + P.o_xin;
      const int r = r0 + 4 * hf + (lane & 3);
      const bool ok = r < valid;
      for (int j = lane >> 2; j < D + C; j += 8) {
        if (j < D) {
          // distrax.ScalarAffine.inverse: (x - shift) * (1/scale)
          xin0[j * RP + r] = ok ? (st[r * D + j] - cst[j]) * cst[D + j] : 0.f;
          ld[j * RP + r] = 0.f;
        } else {
          // Standardizer: (y - mean) / std
          const int c = j - D;
          const float v = ok ? __fdiv_rn(st[P.R * D + r * C + c] - cst[2 * D + c], cst[2 * D + C + c]) : 0.f;
          for (int l = 0; l < L; ++l) sm[P.o_xin + l * P.xin_stride + j * RP + r] = v;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sempty[ti & 1]);
    pair_sync();

    // ---- forward through the L MAF layers
    stamp(3);
    for (int k = 0; k < L; ++k) {
      const float* wf = w_acquire(req);
      stamp(90 + k);
      float* slot = sm + P.o_slot + (BWD ? k : 0) * P.slot_stride;
      const float* xin_k = sm + P.o_xin + k * P.xin_stride + r0;
      float* xnext = ((k + 1 < L) ? sm + P.o_xin + (k + 1) * P.xin_stride : gu) + r0;
      const float* in = xin_k;
#pragma unroll UL
      for (int i = 0; i < n_lin; ++i) {
        const LinDesc& ln = P.lin[i];
        const float* W = wf + ln.w_off;
        const float* bias = wf + ln.b_off;
        int first, nblk;
        my_blocks(A.cf[i], first, nblk);
        if (i < n_lin - 1) {
          float* hout = slot + P.h_off[i + 1] + r0;
          float* dout = slot + P.d_off[i + 1] + r0;
          const int N = ln.N;
          if (has_rows && nblk > 0)
            chain_gemm(in, RP, W, ln.NP, first, nblk, A.cf[i].nsteps, lane, [&](int c, int rr, const float (&v)[2][2]) {
              if (!rows_ok(rr)) return;
#pragma unroll
              for (int cc = 0; cc < 2; ++cc)
                if (c + cc < N) {      // row N of the next operand is its ones row (bias gradient): pad columns stay out
                  const float bj = bias[c + cc];
                  float hv[2], dv[2];
#pragma unroll
                  for (int q = 0; q < 2; ++q) act_eval<RELU>(P.act, v[cc][q] + bj, hv[q], dv[q]);
                  stv(hout + (c + cc) * RP + rr, hv);
                  if (!RELU && BWD) stv(dout + (c + cc) * RP + rr, dv);
                }
            });
          in = hout;
        } else {
          // the last linear's image interleaves (mu_j, logp_j): a lane ends with both; the affine
          // autoregressive transform (jaxili/model.py:739-742) runs right here
          float* sbuf = slot + P.s_off + r0;
          float* vbuf = slot + P.v_off + r0;
          if (has_rows && nblk > 0)
            chain_gemm(in, RP, W, ln.NP, first, nblk, A.cf[i].nsteps, lane, [&](int ce, int rr, const float (&v)[2][2]) {
              const int j = ce >> 1;
              if (j < D && rows_ok(rr)) {
                float bb[2], xv[2], sv[2], vv[2], lv[2];
                ldv(bb, bias + ce);
                ldv(xv, xin_k + j * RP + rr);
                ldv(lv, ld + r0 + j * RP + rr);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  const float lp = v[1][q] + bb[1];
                  sv[q] = expf(0.5f * lp);
                  vv[q] = (xv[q] - (v[0][q] + bb[0])) * sv[q];
                  lv[q] += 0.5f * lp;
                }
                if (BWD) {
                  stv(sbuf + j * RP + rr, sv);
                  stv(vbuf + j * RP + rr, vv);
                }
                stv(ld + r0 + j * RP + rr, lv);
                const int jj = P.reverse ? D - 1 - j : j;
                stv(xnext + jj * RP + rr, vv);
              }
            });
        }
        pair_sync();
        stamp(100 + 10 * k + i);
      }
      w_release(req);
      ++req;
    }

    // ---- log-probability per row (jaxili/model.py:919-921 + 1094-1098); seed of the reverse pass.
    //      This warp's row block, 8 lanes per row stride the features; the last layer's elementwise reverse step
    //      is done here.
    float* gset = sm + P.o_gb + (sq & 1) * P.gb_stride;          // gradient-buffer set of the coming layer step
    if (BWD && sq >= 2) nbar_sync(CH_BAR_EMPTY + (sq & 1), NB);  // its previous content has been consumed
    if (own_rb) {
      const int r = r0 + 4 * hf + (lane >> 3);
      const int f0 = lane & 7;
      const bool vok = r < valid;
      const float* pslot = sm + P.o_slot + (BWD ? L - 1 : 0) * P.slot_stride;
      float* gx_l = gx + ((L - 1) & 1) * P.DP * RP;
      float acc = 0.f, ldsum = 0.f;
      for (int j = f0; j < D; j += 8) {
        const float u = gu[j * RP + r];
        const float l = ld[j * RP + r];
        ldsum += l;
        acc += l - 0.5f * u * u;
        if (BWD) {
          const float gv = vok ? u * A.inv_b : 0.f;
          const int jm = P.reverse ? D - 1 - j : j;      // output j of the flow is MADE feature jm of the last layer
          const float s = pslot[P.s_off + jm * RP + r], v = pslot[P.v_off + jm * RP + r];
          gset[jm * RP + r] = vok ? -gv * s : 0.f;
          gset[(D + jm) * RP + r] = vok ? 0.5f * gv * v - 0.5f * A.inv_b : 0.f;
          gx_l[jm * RP + r] = vok ? gv * s : 0.f;
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        ldsum += __shfl_xor_sync(0xffffffffu, ldsum, o);
      }
      const float lp = acc - (float)D * kHalfLog2Pi + A.logdet_const;
      const bool lead = f0 == 0 && vok;
      if (lead && A.out_logp) A.out_logp[row0 + r] = lp;
      if (!BWD && lead && A.out_logdet) A.out_logdet[row0 + r] = ldsum;
      if (BWD) my_loss += warp_sum(lead ? -lp * A.inv_b : 0.f);
      if (!BWD && A.out_u) {   // own rows are contiguous in HBM
        const int rw = r0 + 4 * hf;
        const int nrow = min(4, valid - rw);
        float* dst = A.out_u + (row0 + rw) * D;
        for (int idx = lane; idx < nrow * D; idx += 32) {
          const int rr = idx / D, j = idx - rr * D;
          dst[idx] = gu[j * RP + rw + rr];
        }
      }
    }
    if (BWD) pair_sync();
    stamp(4);

    // ---- reverse pass
    if (BWD) {
      for (int k = L - 1; k >= 0; --k, ++sq) {
        const float* wb = w_acquire(req);
        stamp(190 + k);
        const float* slot = sm + P.o_slot + k * P.slot_stride;
        const float* gx_k = gx + (k & 1) * P.DP * RP + r0;
        gset = sm + P.o_gb + (sq & 1) * P.gb_stride;
        float* gnext = sm + P.o_gb + ((sq + 1) & 1) * P.gb_stride;   // set of the next layer step (its obuf is written here)
#pragma unroll UL
        for (int ii = 0; ii < n_lin; ++ii) {
          const int i = n_lin - 1 - ii;
          const LinDesc& ln = P.lin[i];
          if (i == 0) {
            // every gradient of this layer step is in its buffer set: hand it to the gradient warps
            nbar_arrive(CH_BAR_FULL + (sq & 1), NB);
            stamp(200 + 10 * k);
            if (k == 0) break;                          // nothing consumes the gradient w.r.t. the flow's input
            if (sq + 1 >= 2) nbar_sync(CH_BAR_EMPTY + ((sq + 1) & 1), NB);   // next set: previous content consumed
            stamp(300 + 10 * k);
          }
          int first, nblk;
          my_blocks(A.cx[i], first, nblk);
          if (has_rows && nblk > 0) {
            const float* G = gset + P.g_off[i] + r0;
            const float* Wt = wb + ln.wt_off;
            if (i > 0) {
              // through the activation: g_a = g_h * act'(a)
              const float* dsrc = slot + (RELU ? P.h_off[i] : P.d_off[i]) + r0;
              float* gout = gset + P.g_off[i - 1] + r0;
              chain_gemm(G, RP, Wt, ln.KP, first, nblk, A.cx[i].nsteps, lane, [&](int c, int rr, const float (&v)[2][2]) {
                if (!rows_ok(rr)) return;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                  float dv[2], gv[2];
                  ldv(dv, dsrc + (c + cc) * RP + rr);
#pragma unroll
                  for (int q = 0; q < 2; ++q) gv[q] = RELU ? (dv[q] > 0.f ? v[cc][q] : 0.f) : v[cc][q] * dv[q];
                  stv(gout + (c + cc) * RP + rr, gv);
                }
              });
            } else {
              // gradient w.r.t. this layer's input x_k (MADE path + direct path); x_k is layer k-1's flipped
              // output: continue straight into that layer's elementwise reverse step
              const float* pslot = sm + P.o_slot + (k - 1) * P.slot_stride + r0;
              float* gx_p = gx + ((k - 1) & 1) * P.DP * RP + r0;
              float* onext = gnext + r0;
              chain_gemm(G, RP, Wt, ln.KP, first, nblk, A.cx[i].nsteps, lane, [&](int c, int rr, const float (&v)[2][2]) {
                if (!rows_ok(rr)) return;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
                  if (c + cc < D) {
                    float dv[2], sv[2], vv[2], o0[2], o1[2], gn[2];
                    ldv(dv, gx_k + (c + cc) * RP + rr);
                    const int j = P.reverse ? D - 1 - (c + cc) : (c + cc);
                    ldv(sv, pslot + P.s_off + j * RP + rr);
                    ldv(vv, pslot + P.v_off + j * RP + rr);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                      const float gv = v[cc][q] + dv[q];
                      const bool okr = r0 + rr + q < valid;
                      o0[q] = okr ? -gv * sv[q] : 0.f;
                      o1[q] = okr ? 0.5f * gv * vv[q] - 0.5f * A.inv_b : 0.f;
                      gn[q] = okr ? gv * sv[q] : 0.f;
                    }
                    stv(onext + j * RP + rr, o0);
                    stv(onext + (D + j) * RP + rr, o1);
                    stv(gx_p + j * RP + rr, gn);
                  }
              });
            }
          }
          pair_sync();
          stamp(400 + 10 * k + i);
        }
        w_release(req);
        ++req;
      }
    }
  }

  if (BWD) {   // per-warp loss sums in warp order: reproducible
    if (lane == 0) loss_red[cw] = my_loss;
    nbar_sync(CH_BAR_CHAIN, 32 * CH_CW);
    if (cw == 0 && lane == 0) {
      float tsum = 0.f;
      for (int w = 0; w < CH_CW; ++w) tsum += loss_red[w];
      A.partial_loss[blockIdx.x] = tsum;
    }
  }
  stamp(99);
  stamp_end();
#ifdef MAFB_CTA_TIMES
  if (A.trace && cw == 0 && lane == 0) {
    unsigned long long tns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tns));
    A.trace[400 + blockIdx.x] = (long long)tns;
  }
#endif
}

}  // namespace mafb
