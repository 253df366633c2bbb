// chain_kernel.cuh -- warp-specialised fused ConditionalMAF kernel for sm_100a (the default flow kernel).
//
// Same contract, shared-memory plan and weight images as flow_kernel.cuh (forward: jaxili/model.py:1092-1098,
// 903-921, 848-874, 718-743, 665-686, 517-544; reverse: what jax.value_and_grad produces for loss.py:35-83),
// different execution structure.  At the batch sizes of the narrow configs an SM sees ~28 rows, and a MADE is a
// chain of 15 forward + 15 reverse dependent linears: a design that spreads every linear over the whole CTA pays
// a CTA barrier, a pipeline fill and a drain per linear (47 % barrier stall in round 1).  Here
//
//   * CHAIN warps own rows.  Warp c holds rows 4c..4c+3 through every linear of every layer, forward and
//     reverse: the only synchronisation on the dependent chain is __syncwarp.  Two chain warps per SM
//     sub-partition, issued with priority (highest warp ids).  A lane holds 4 rows x 4 (or 8) columns and one of
//     4 (or 8) parts of the reduction; packed FFMA2; operands prefetched through a 4-stage register ring; the
//     parts are combined by a shuffle reduce-scatter that leaves every lane 2 columns x 2 rows for the epilogue.
//   * GRADIENT warps compute the weight/bias gradients [A;1]^T G, the only product that crosses rows.  They
//     consume the per-linear gradient buffers asynchronously (named barriers full[i] / empty[i]) and fill the
//     issue slots the chain leaves empty; they never sit on the critical path.
//   * one PRODUCER warp streams the weight images and the input tiles with bulk copies (TMA 1-D) behind
//     full/empty mbarriers.
#pragma once
#include "common.cuh"
#include "flow_kernel.cuh"

namespace mafb {

constexpr int CH_CW = 8;                         // chain warps: 4 rows each, 32-row tiles
constexpr int CH_GW = 7;                         // gradient warps (16 warps per CTA: 128 registers per thread)
constexpr int CH_NT_BWD = 32 * (1 + CH_GW + CH_CW);
constexpr int CH_NT_FWD = 32 * (1 + CH_CW);
constexpr int CH_MAX_LIN = 6;
constexpr int CH_BAR_FULL = 1;                   // named barriers: full[i] = 1 + i, empty[i] = 7 + i
constexpr int CH_BAR_EMPTY = 7;
constexpr int CH_BAR_TILE = 13;                  // gradient warps are done with the tile's stash
constexpr int CH_BAR_CHAIN = 14;                 // chain warps only (loss hand-off at the end)

__device__ __forceinline__ void nbar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void nbar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// acc[j][i] += a[i] * w[j]: 8 packed FFMA2 (row pairs, the weight as the broadcast operand)
__device__ __forceinline__ void ch_fma_tile(float (&acc)[4][4], const float (&a)[4], const float (&w)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; i += 2) ffma2(acc[j][i], acc[j][i + 1], a[i], a[i + 1], w[j], w[j]);
}

// n reduction steps of G 4x4 tiles that share the row operand; pa / pw advance by da / dw floats per step.
// n % 4 steps are peeled off the front; the rest runs through a 4-stage register ring: the loads of step s+4 are
// issued behind the FMAs of step s, i.e. 3 steps (>= 48 FMA-pipe cycles) ahead of their use -- more than the
// 29-cycle shared-memory latency.
template <int G>
__device__ __forceinline__ void ch_mac(float (&acc)[G][4][4], const float* __restrict__ pa,
                                       const float* const (&pw0)[G], int n, int da, int dw) {
  const float* pw[G];
#pragma unroll
  for (int g = 0; g < G; ++g) pw[g] = pw0[g];
  const int rem = n & 3;
#pragma unroll 1
  for (int u = 0; u < rem; ++u) {
    float a[4], w[G][4];
    ldv(a, pa);
#pragma unroll
    for (int g = 0; g < G; ++g) ldv(w[g], pw[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) ch_fma_tile(acc[g], a, w[g]);
    pa += da;
#pragma unroll
    for (int g = 0; g < G; ++g) pw[g] += dw;
  }
  int m = n - rem;                 // multiple of 4
  if (m == 0) return;
  float a[4][4], w[4][G][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    ldv(a[u], pa + u * da);
#pragma unroll
    for (int g = 0; g < G; ++g) ldv(w[u][g], pw[g] + u * dw);
  }
#pragma unroll 1
  for (m -= 4; m > 0; m -= 4) {
    pa += 4 * da;
#pragma unroll
    for (int g = 0; g < G; ++g) pw[g] += 4 * dw;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int g = 0; g < G; ++g) ch_fma_tile(acc[g], a[u], w[u][g]);
      ldv(a[u], pa + u * da);
#pragma unroll
      for (int g = 0; g < G; ++g) ldv(w[u][g], pw[g] + u * dw);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int g = 0; g < G; ++g) ch_fma_tile(acc[g], a[u], w[u][g]);
}

// Reduce-scatter of a 4x4 tile (acc[col][row]) over the lanes that hold its reduction parts.
// stage 1 (lane bit 4): keep column pair b1;  stage 2 (lane bit 3): keep row pair b0.
__device__ __forceinline__ void ch_rs_2x2(const float (&acc)[4][4], bool b1, bool b0, float (&r)[2][2]) {
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
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = b0 ? t[c][i] : t[c][2 + i];
      const float mine = b0 ? t[c][2 + i] : t[c][i];
      r[c][i] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
    }
}

// One product of the chain for the warp's 4 rows: out[c][r] = sum_q in[q][r] * W[q][c], all column blocks.
//   KS = 4: lane = kq * 8 + cbl, the lane's reduction part is q = kq, kq+4, ...; G column blocks per lane
//           (cbl and cbl + 8).  After the reduce-scatter a lane owns columns (ce, ce+1) x rows (re, re+1):
//           epi(ce, re, v) with v[col][row].
//   KS = 8: lane = kq * 4 + cbl, G = 1; a third stage (lane bit 2) leaves one column: epi(c, re, v) uses v[0][*].
struct NoStamp { __device__ __forceinline__ void operator()(int) const {} };
template <int KS, int G, class Epi, class St = NoStamp>
__device__ __forceinline__ void chain_gemm(const float* __restrict__ in, int RP, const float* __restrict__ W, int ldw,
                                           int nCB, int nsteps, int n_pass, int lane, Epi&& epi, St&& st = St()) {
  static_assert((KS == 4 && (G == 1 || G == 2)) || (KS == 8 && G == 1), "supported lane maps");
  constexpr int CPL = 32 / KS;
  const int cbl = lane & (CPL - 1);
  const int kq = lane / CPL;
  const bool b1 = (lane & 16) != 0, b0 = (lane & 8) != 0;
#pragma unroll 1
  for (int pass = 0; pass < n_pass; ++pass) {
    float acc[G][4][4];
    const float* pw[G];
    bool ok[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int cb = (pass * G + g) * CPL + cbl;
      ok[g] = cb < nCB;
      pw[g] = W + kq * ldw + 4 * (ok[g] ? cb : nCB - 1);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[g][j][i] = 0.f;
    }
    st(1);
    ch_mac<G>(acc, in + kq * RP, pw, nsteps, KS * RP, KS * ldw);
    st(2);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int c0 = 4 * ((pass * G + g) * CPL + cbl);
      float r[2][2];
      ch_rs_2x2(acc[g], b1, b0, r);
      st(3 + g);
      if (KS == 4) {
        if (ok[g]) epi(c0 + (b1 ? 2 : 0), b0 ? 2 : 0, r);
      } else {
        const bool b2 = (lane & 4) != 0;   // stage 3: keep column b2 of the pair
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = b2 ? r[0][i] : r[1][i];
          const float mine = b2 ? r[1][i] : r[0][i];
          r[0][i] = mine + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        if (ok[g]) epi(c0 + (b1 ? 2 : 0) + (b2 ? 1 : 0), b0 ? 2 : 0, r);
      }
    }
  }
}

// dispatch on the host-chosen lane map; epi1(c, re, v[2 rows]) is called once per finished column
template <class Epi, class St = NoStamp>
__device__ __forceinline__ void chain_gemm_cols(const float* __restrict__ in, int RP, const float* __restrict__ W,
                                                int ldw, const FlowArgs::ChainPh& ph, int lane, Epi&& epi1, St&& st = St()) {
  auto both = [&](int ce, int re, const float (&v)[2][2]) {
    epi1(ce, re, v[0]);
    epi1(ce + 1, re, v[1]);
  };
  auto one = [&](int c, int re, const float (&v)[2][2]) { epi1(c, re, v[0]); };
  if (ph.ks_shift == 3) chain_gemm<8, 1>(in, RP, W, ldw, ph.nCB, ph.nsteps, ph.n_pass, lane, one);
  else if (ph.groups == 2) chain_gemm<4, 2>(in, RP, W, ldw, ph.nCB, ph.nsteps, ph.n_pass, lane, both, st);
  else chain_gemm<4, 1>(in, RP, W, ldw, ph.nCB, ph.nsteps, ph.n_pass, lane, both, st);
}
// KS = 4 only: epi(ce, re, v[2 cols][2 rows]) gets the aligned column pair (ce even)
template <class Epi>
__device__ __forceinline__ void chain_gemm_pairs(const float* __restrict__ in, int RP, const float* __restrict__ W,
                                                 int ldw, const FlowArgs::ChainPh& ph, int lane, Epi&& epi) {
  if (ph.groups == 2) chain_gemm<4, 2>(in, RP, W, ldw, ph.nCB, ph.nsteps, ph.n_pass, lane, epi);
  else chain_gemm<4, 1>(in, RP, W, ldw, ph.nCB, ph.nsteps, ph.n_pass, lane, epi);
}

// NLIN > 0 fixes the number of masked linears per MADE at compile time (the loops over them unroll and the
// per-linear descriptors become constant-bank operands); NLIN == 0 reads it from the plan.
template <bool BWD, bool RELU, int NLIN>
__global__ void __launch_bounds__(BWD ? CH_NT_BWD : CH_NT_FWD, 1)
maf_chain_kernel(const __grid_constant__ Plan P, const __grid_constant__ FlowArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* wfull = reinterpret_cast<uint64_t*>(smem_raw);   // [2] weight image landed
  uint64_t* wempty = wfull + 2;                              // [2] all chain warps are done with it
  uint64_t* sfull = wfull + 4;                               // [2] input tile landed
  uint64_t* sempty = wfull + 6;                              // [2] all chain warps have transposed it
  float* loss_red = reinterpret_cast<float*>(smem_raw + 64); // [CH_CW]
  float* sm = reinterpret_cast<float*>(smem_raw + MAFB_SMEM_HDR);

  constexpr int NT = BWD ? CH_NT_BWD : CH_NT_FWD;
  constexpr int GW = BWD ? CH_GW : 0;
  constexpr int NB = 32 * (GW + CH_CW);                      // threads on the full / empty / tile barriers
  constexpr int UL = NLIN > 0 ? NLIN : 1;                    // unroll factor of the loops over the linears
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = P.D, C = P.C, L = P.L, RP = P.RP;
  const int n_lin = NLIN > 0 ? NLIN : P.n_lin;
  const int Rt = A.Rt, nRB = Rt / 4;
  const long long row_base = A.batch_index ? (long long)(A.batch_index[0] % A.n_batches) * A.B : 0;
  const int grid = gridDim.x;
  const int my_n = (A.n_tiles - (int)blockIdx.x + grid - 1) / grid;
  const int per_tile = BWD ? 2 * L : L;
  const int total_req = my_n * per_tile;
  const bool two_wbuf = P.n_wbuf == 2;

#ifdef MAFB_TRACE
  // CTA 0: chain warp 0 stamps into trace[0..499], gradient warp 0 into trace[500..999] ((id, clock) pairs)
  int n_trace = 0;
  const bool tr_on = A.trace && blockIdx.x == 0 && lane == 0 && (warp == GW + 1 || (BWD && warp == 1));
  long long* const tbase = A.trace + ((BWD && warp == 1) ? 500 : 0);
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
  __syncthreads();
  float* const cst = sm + P.o_cst;   // [shift D][inv_scale D][mean C][std C]
  for (int idx = tid; idx < 2 * D + 2 * C; idx += NT) cst[idx] = A.stdc[idx];
  if (BWD) {
    // a feature row of ones behind every weight-gradient A operand: [A;1]^T G yields the bias gradient
    for (int idx = tid; idx < RP; idx += NT) {
      for (int l = 0; l < L; ++l) {
        sm[P.o_xin + l * P.xin_stride + P.lin[0].KP * RP + idx] = 1.f;
        for (int i = 1; i < n_lin; ++i)
          sm[P.o_slot + l * P.slot_stride + P.h_off[i] + P.lin[i].KP * RP + idx] = 1.f;
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
        if (use > 0) mbar_wait(&wempty[buf], (uint32_t)((use - 1) & 1));
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
        if (lane == 0) mbar_wait(&sempty[ti & 1], (uint32_t)((ti >> 1) & 1));
        __syncwarp();
        stage_issue(ti + 2);
      }
      if (++j == per_tile) { j = 0; ++ti; }
    }
    return;
  }

  // ======================================================================== gradient warps
  if (BWD && warp <= GW) {
    const int gw = warp - 1;
    float* pg_cta = A.partial_grad + (size_t)blockIdx.x * P.P;
    for (int ti = 0; ti < my_n; ++ti) {
      const bool accum = ti > 0;
      const bool last_tile = ti + 1 == my_n;
      for (int k = L - 1; k >= 0; --k) {
        float* pg = pg_cta + (size_t)k * P.ppl;
        const float* xin = sm + P.o_xin + k * P.xin_stride;
        const float* slot = sm + P.o_slot + k * P.slot_stride;
#pragma unroll UL
        for (int ii = 0; ii < n_lin; ++ii) {
          const int i = n_lin - 1 - ii;
          const LinDesc& ln = P.lin[i];
          const FlowArgs::Phase& ph = A.ph[i];
          const float* G = sm + P.o_gl[i];
          const float* Ain = (i == 0) ? xin : slot + P.h_off[i];
          nbar_sync(CH_BAR_FULL + i, NB);
          stamp(500 + 10 * k + i);
          for (int blk = gw; blk < ph.bB; blk += GW) {
            int b = blk, g = 0;
            while (b >= ph.nKG) { b -= ph.nKG; ++g; }
            const int l16 = lane & 15;
            const int kb = b * (16 >> ph.nb_shift) + (l16 >> ph.nb_shift);
            const int nb = (g << ph.nb_shift) + (l16 & ((1 << ph.nb_shift) - 1));
            mm_inner_rs2(Ain, ln.K, ln.KP, G, ln.N, Rt, RP, kb, ph.nKB, nb < ph.nNB ? nb : ph.nNB - 1, ph.nNB,
                         lane >> 4, nb < ph.nNB && kb < ph.nKB, pg + ln.pw_off, accum);
            if (k == 3) stamp(800 + i);
          }
          stamp(600 + 10 * k + i);
          if (!(last_tile && k == 0)) nbar_arrive(CH_BAR_EMPTY + i, NB);
        }
      }
      if (!last_tile) nbar_arrive(CH_BAR_TILE, NB);
    }
    stamp_end();
    return;
  }

  // ======================================================================== chain warps
  const int cw = warp - GW - 1;
  const bool has_rows = cw < nRB;
  const int r0 = 4 * (has_rows ? cw : 0);
  float* const obuf = sm + P.o_obuf;
  float* const gu = sm + P.o_gu;
  float* const gx = sm + P.o_gx;
  float* const ld = sm + P.o_ld;
  float my_loss = 0.f;
  int req = 0;

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

  for (int ti = 0; ti < my_n; ++ti) {
    const int t = blockIdx.x + ti * grid;
    const long long row0 = (long long)t * Rt;
    const int valid = (int)min((long long)Rt, A.B - row0);

    // the previous tile's activations are still being read by the gradient warps
    if (BWD && ti > 0) nbar_sync(CH_BAR_TILE, NB);

    // ---- stage in: own rows, raw [row][feature] -> standardised feature-major
    stamp(1);
    mbar_wait(&sfull[ti & 1], (uint32_t)((ti >> 1) & 1));
    stamp(2);
    if (has_rows) {
      const float* st = sm + P.o_stage + (ti & 1) * P.stage_stride;
      float* xin0 = sm + P.o_xin;
      const int r = r0 + (lane & 3);
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

    // ---- forward through the L MAF layers
    stamp(3);
    for (int k = 0; k < L; ++k) {
      const float* wf = w_acquire(req);
      stamp(90 + k);
      if (has_rows) {
        float* slot = sm + P.o_slot + (BWD ? k : 0) * P.slot_stride;
        const float* xin_k = sm + P.o_xin + k * P.xin_stride;
        float* xnext = (k + 1 < L) ? sm + P.o_xin + (k + 1) * P.xin_stride : gu;
        const float* in = xin_k + r0;
#pragma unroll UL
        for (int i = 0; i < n_lin; ++i) {
          const LinDesc& ln = P.lin[i];
          const float* W = wf + ln.w_off;
          const float* bias = wf + ln.b_off;
          if (i < n_lin - 1) {
            float* hout = slot + P.h_off[i + 1] + r0;
            float* dout = slot + P.d_off[i + 1] + r0;
            chain_gemm_cols(in, RP, W, ln.NP, A.cf[i], lane, [&](int c, int re, const float (&v)[2]) {
              const float bj = bias[c];
              float hv[2], dv[2];
#pragma unroll
              for (int q = 0; q < 2; ++q) act_eval<RELU>(P.act, v[q] + bj, hv[q], dv[q]);
              stv(hout + c * RP + re, hv);
              if (!RELU && BWD) stv(dout + c * RP + re, dv);
            }, [&](int id) { if (k == 1) stamp(700 + 10 * i + id); });
            in = hout;
          } else {
            // the last linear's image interleaves (mu_j, logp_j): a lane ends with both for 2 rows; the affine
            // autoregressive transform (jaxili/model.py:739-742) runs right here
            float* sbuf = slot + P.s_off + r0;
            float* vbuf = slot + P.v_off + r0;
            chain_gemm_pairs(in, RP, W, ln.NP, A.cf[i], lane, [&](int ce, int re, const float (&v)[2][2]) {
              const int j = ce >> 1;
              if (j < D) {
                float bb[2], xv[2], sv[2], vv[2], lv[2];
                ldv(bb, bias + ce);
                ldv(xv, xin_k + r0 + j * RP + re);
                ldv(lv, ld + r0 + j * RP + re);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  const float lp = v[1][q] + bb[1];
                  sv[q] = expf(0.5f * lp);
                  vv[q] = (xv[q] - (v[0][q] + bb[0])) * sv[q];
                  lv[q] += 0.5f * lp;
                }
                if (BWD) {
                  stv(sbuf + j * RP + re, sv);
                  stv(vbuf + j * RP + re, vv);
                }
                stv(ld + r0 + j * RP + re, lv);
                const int jj = P.reverse ? D - 1 - j : j;
                stv(xnext + r0 + jj * RP + re, vv);
              }
            });
          }
          __syncwarp();
          stamp(100 + 10 * k + i);
        }
      }
      w_release(req);
      ++req;
    }

    // ---- log-probability per row (jaxili/model.py:919-921 + 1094-1098); seed of the reverse pass.
    //      8 lanes per row stride the features; the last layer's elementwise reverse step is done here.
    if (BWD && ti > 0) nbar_sync(CH_BAR_EMPTY + n_lin - 1, NB);   // obuf: previous content consumed
    if (has_rows) {
      const int r = r0 + (lane >> 3);
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
          obuf[jm * RP + r] = vok ? -gv * s : 0.f;
          obuf[(D + jm) * RP + r] = vok ? 0.5f * gv * v - 0.5f * A.inv_b : 0.f;
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
        const int nrow = min(4, valid - r0);
        float* dst = A.out_u + (row0 + r0) * D;
        for (int idx = lane; idx < nrow * D; idx += 32) {
          const int rr = idx / D, j = idx - rr * D;
          dst[idx] = gu[j * RP + r0 + rr];
        }
      }
    }
    if (BWD) {
      __syncwarp();
      nbar_arrive(CH_BAR_FULL + n_lin - 1, NB);
    }
    stamp(4);

    // ---- reverse pass
    if (BWD) {
      for (int k = L - 1; k >= 0; --k) {
        const float* wb = w_acquire(req);
        stamp(190 + k);
        const float* slot = sm + P.o_slot + k * P.slot_stride;
        const float* gx_k = gx + (k & 1) * P.DP * RP + r0;
#pragma unroll UL
        for (int ii = 0; ii < n_lin; ++ii) {
          const int i = n_lin - 1 - ii;
          if (i == 0 && k == 0) break;                  // nothing consumes the gradient w.r.t. the flow's input
          const LinDesc& ln = P.lin[i];
          const int tgt = i > 0 ? i - 1 : n_lin - 1;    // buffer this product's epilogue writes
          const bool first_write = i > 0 && ti == 0 && k == L - 1;
          stamp(200 + 10 * k + i);
          if (!first_write) nbar_sync(CH_BAR_EMPTY + tgt, NB);
          stamp(300 + 10 * k + i);
          if (has_rows) {
            const float* G = sm + P.o_gl[i] + r0;
            const float* Wt = wb + ln.wt_off;
            if (i > 0) {
              // through the activation: g_a = g_h * act'(a)
              const float* dsrc = slot + (RELU ? P.h_off[i] : P.d_off[i]) + r0;
              float* gout = sm + P.o_gl[i - 1] + r0;
              chain_gemm_cols(G, RP, Wt, ln.KP, A.cx[i], lane, [&](int c, int re, const float (&v)[2]) {
                float dv[2], gv[2];
                ldv(dv, dsrc + c * RP + re);
#pragma unroll
                for (int q = 0; q < 2; ++q) gv[q] = RELU ? (dv[q] > 0.f ? v[q] : 0.f) : v[q] * dv[q];
                stv(gout + c * RP + re, gv);
              });
            } else {
              // gradient w.r.t. this layer's input x_k (MADE path + direct path); x_k is layer k-1's flipped
              // output: continue straight into that layer's elementwise reverse step
              const float* pslot = sm + P.o_slot + (k - 1) * P.slot_stride + r0;
              float* gx_p = gx + ((k - 1) & 1) * P.DP * RP + r0;
              chain_gemm_cols(G, RP, Wt, ln.KP, A.cx[i], lane, [&](int c, int re, const float (&v)[2]) {
                if (c < D) {
                  float dv[2], sv[2], vv[2], o0[2], o1[2], gn[2];
                  ldv(dv, gx_k + c * RP + re);
                  const int j = P.reverse ? D - 1 - c : c;
                  ldv(sv, pslot + P.s_off + j * RP + re);
                  ldv(vv, pslot + P.v_off + j * RP + re);
#pragma unroll
                  for (int q = 0; q < 2; ++q) {
                    const float gv = v[q] + dv[q];
                    const bool okr = r0 + re + q < valid;
                    o0[q] = okr ? -gv * sv[q] : 0.f;
                    o1[q] = okr ? 0.5f * gv * vv[q] - 0.5f * A.inv_b : 0.f;
                    gn[q] = okr ? gv * sv[q] : 0.f;
                  }
                  stv(obuf + r0 + j * RP + re, o0);
                  stv(obuf + r0 + (D + j) * RP + re, o1);
                  stv(gx_p + j * RP + re, gn);
                }
              });
            }
          }
          __syncwarp();
          nbar_arrive(CH_BAR_FULL + tgt, NB);
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
}

}  // namespace mafb
