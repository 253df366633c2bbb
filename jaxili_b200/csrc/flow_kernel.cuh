// flow_kernel.cuh -- fused whole-flow ConditionalMAF kernel for sm_100a.
//
// One persistent CTA per SM walks row tiles of the batch.  For each tile it keeps, in shared
// memory and registers only: the standardised inputs, every MADE activation of every MAF
// layer, the running log|det J|, and (training) the whole reverse pass.  HBM sees the batch
// rows once (bulk/TMA copy), the log-probabilities or the per-CTA partial gradients.
//
//   forward  (jaxili/model.py:1092-1098, 903-921, 848-874, 718-743, 665-686, 517-544)
//   backward (what jax.value_and_grad produces for jaxili/loss.py:35-83, train.py:330-332)
//
// Per MAF layer the mask-folded weights arrive as one bulk copy of a pre-packed image
// (forward part: W[K][NP]+bias; backward part: W^T[N][KP]) double-buffered behind mbarriers.
#pragma once
#include "common.cuh"

namespace mafb {

struct FlowArgs {
  const float* wimg;          // [L][img_floats] packed images
  const float* x;             // [rows][D]
  const float* y;             // [rows][C]
  long long B;                // rows in this batch
  const int* batch_index;     // optional device scalar selecting the batch
  long long n_batches;
  const float* stdc;          // [shift D][inv_scale D][mean C][std C]
  float logdet_const;         // sum_i -log|scale_i|
  float inv_b;                // 1 / global batch size
  float* out_logp;            // forward-only: [B]
  float* partial_grad;        // training: [grid][P]
  float* partial_loss;        // training: [grid]
  int Rt;                     // rows per tile in this launch (multiple of 4, <= Plan::R)
  int n_tiles;
};

constexpr int FM = 4, FN = 2;   // forward tile: rows x cols per thread
constexpr int XM = 4, XN = 4;   // input-gradient tile
constexpr int BK = 4, BN = 4;   // weight-gradient tile
constexpr float kHalfLog2Pi = 0.9189385332046727f;

template <bool BWD, bool RELU, int NT>
__global__ void __launch_bounds__(NT, 1)
maf_flow_kernel(const __grid_constant__ Plan P, const __grid_constant__ FlowArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem_raw);   // [2] weight images
  uint64_t* sbar = wbar + 2;                                // [2] input tiles
  float* loss_acc = reinterpret_cast<float*>(smem_raw + 32);
  float* sm = reinterpret_cast<float*>(smem_raw + 64);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const int D = P.D, C = P.C, L = P.L, RP = P.RP, n_lin = P.n_lin;
  const int Rt = A.Rt;
  const int nRB = Rt / FM;
  const long long row_base =
      A.batch_index ? (long long)(A.batch_index[0] % A.n_batches) * A.B : 0;
  const int grid = gridDim.x;
  const int my_n = (A.n_tiles - (int)blockIdx.x + grid - 1) / grid;

  if (tid == 0) {
    mbar_init(&wbar[0], 1);
    mbar_init(&wbar[1], 1);
    mbar_init(&sbar[0], 1);
    mbar_init(&sbar[1], 1);
    *loss_acc = 0.f;
    mbar_fence_init();
  }
  __syncthreads();

  const bool recompute = BWD && !P.stash_all;
  const int per_tile = BWD ? 2 * L : L;
  const int total_req = my_n * per_tile;

  // ---- weight image requests: s-th request of this CTA (thread 0 only)
  auto w_issue = [&](int s) {
    const int j = s % per_tile;
    int layer, part;
    if (!BWD || j < L) { layer = j; part = 0; } else { layer = 2 * L - 1 - j; part = 1; }
    const float* src = A.wimg + (size_t)layer * P.img_floats;
    uint32_t bytes;
    if (part == 0) bytes = P.fwd_floats * 4u;
    else if (recompute) bytes = P.img_floats * 4u;
    else { src += P.fwd_floats; bytes = P.bwd_floats * 4u; }
    const int buf = s % P.n_wbuf;
    fence_proxy_async();
    mbar_expect_tx(&wbar[buf], bytes);
    bulk_g2s(sm + P.o_wbuf + buf * P.wbuf_stride, src, bytes, &wbar[buf]);
  };
  // every thread: make request s resident, prefetch s+1; returns the buffer
  auto w_acquire = [&](int s) -> const float* {
    const int buf = s % P.n_wbuf;
    if (tid == 0) {
      if (P.n_wbuf == 1) w_issue(s);
      else if (s + 1 < total_req) w_issue(s + 1);
    }
    mbar_wait(&wbar[buf], (uint32_t)((s / P.n_wbuf) & 1));
    return sm + P.o_wbuf + buf * P.wbuf_stride;
  };

  // ---- input tile staging: all threads call (uniform)
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
      if (tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(&sbar[i & 1], (uint32_t)(valid * (D + C) * 4));
        bulk_g2s(dst, xs, (uint32_t)(valid * D * 4), &sbar[i & 1]);
        if (C) bulk_g2s(dst + P.R * D, ys, (uint32_t)(valid * C * 4), &sbar[i & 1]);
      }
    } else {
      for (int idx = tid; idx < valid * D; idx += NT) dst[idx] = xs[idx];
      for (int idx = tid; idx < valid * C; idx += NT) dst[P.R * D + idx] = ys[idx];
      if (tid == 0) mbar_arrive(&sbar[i & 1]);
    }
  };

  if (tid == 0 && P.n_wbuf == 2 && total_req > 0) w_issue(0);
  if (my_n > 0) stage_issue(0);

  float* const obuf = sm + P.o_obuf;
  float* const gu = sm + P.o_gu;
  float* const gx = sm + P.o_gx;
  float* const ld = sm + P.o_ld;
  const float* const c_shift = A.stdc;
  const float* const c_iscale = A.stdc + D;
  const float* const c_mean = A.stdc + 2 * D;
  const float* const c_std = A.stdc + 2 * D + C;

  // ---- one MADE + affine transform.  replay: recomputation inside the reverse pass
  //      (no log-det accumulation, no hand-off of v to the next layer).
  auto layer_forward = [&](int k, const float* wf, float* slot, bool replay) {
    const float* in = sm + P.o_xin + k * P.xin_stride;
    for (int i = 0; i < n_lin; ++i) {
      const LinDesc& ln = P.lin[i];
      const float* W = wf + ln.w_off;
      const float* bias = wf + ln.b_off;
      const int items = nRB * (ln.NP / FN);
      if (i < n_lin - 1) {
        float* hout = slot + P.h_off[i + 1];
        float* dout = slot + P.d_off[i + 1];
        for (int it = tid; it < items; it += NT) {
          const int r0 = (it % nRB) * FM, c0 = (it / nRB) * FN;
          mm_outer<FM, FN>(in, ln.K, RP, W, ln.NP, r0, c0, [&](float (&acc)[FN][FM]) {
#pragma unroll
            for (int j = 0; j < FN; ++j) {
              const float bj = bias[c0 + j];
              float hv[FM], dv[FM];
#pragma unroll
              for (int q = 0; q < FM; ++q) act_eval<RELU>(P.act, acc[j][q] + bj, hv[q], dv[q]);
              stv(hout + (c0 + j) * RP + r0, hv);
              if (!RELU && BWD) stv(dout + (c0 + j) * RP + r0, dv);
            }
          });
        }
        in = hout;
      } else {
        for (int it = tid; it < items; it += NT) {
          const int r0 = (it % nRB) * FM, c0 = (it / nRB) * FN;
          mm_outer<FM, FN>(in, ln.K, RP, W, ln.NP, r0, c0, [&](float (&acc)[FN][FM]) {
#pragma unroll
            for (int j = 0; j < FN; ++j) {
              const float bj = bias[c0 + j];
              float ov[FM];
#pragma unroll
              for (int q = 0; q < FM; ++q) ov[q] = acc[j][q] + bj;
              stv(obuf + (c0 + j) * RP + r0, ov);
            }
          });
        }
      }
      __syncthreads();
    }
    // affine autoregressive transform (jaxili/model.py:739-742)
    const float* xin = sm + P.o_xin + k * P.xin_stride;
    float* xnext = (k + 1 < L) ? sm + P.o_xin + (k + 1) * P.xin_stride : gu;
    float* sbuf = slot + P.s_off;
    float* vbuf = slot + P.v_off;
    for (int idx = tid; idx < D * Rt; idx += NT) {
      const int j = idx / Rt, r = idx - j * Rt;
      const float mu = obuf[j * RP + r], lp = obuf[(D + j) * RP + r];
      const float s = expf(0.5f * lp);
      const float v = (xin[j * RP + r] - mu) * s;
      if (BWD) { sbuf[j * RP + r] = s; vbuf[j * RP + r] = v; }
      if (!replay) {
        ld[j * RP + r] += 0.5f * lp;
        const int jj = P.reverse ? D - 1 - j : j;
        xnext[jj * RP + r] = v;
      }
    }
    __syncthreads();
  };

  for (int ti = 0; ti < my_n; ++ti) {
    const int t = blockIdx.x + ti * grid;
    const long long row0 = (long long)t * Rt;
    const int valid = (int)min((long long)Rt, A.B - row0);
    int req = ti * per_tile;

    // ---- stage in: wait for the bulk copy, transpose to feature-major, standardise
    mbar_wait(&sbar[ti & 1], (uint32_t)((ti >> 1) & 1));
    __syncthreads();
    {
      const float* st = sm + P.o_stage + (ti & 1) * P.stage_stride;
      float* xin0 = sm + P.o_xin;
      for (int idx = tid; idx < D * Rt; idx += NT) {
        const int j = idx / Rt, r = idx - j * Rt;
        // distrax.ScalarAffine.inverse: (x - shift) * (1/scale)
        xin0[j * RP + r] = r < valid ? (st[r * D + j] - c_shift[j]) * c_iscale[j] : 0.f;
      }
      for (int idx = tid; idx < C * Rt; idx += NT) {
        const int c = idx / Rt, r = idx - c * Rt;
        // Standardizer: (y - mean) / std
        const float v = r < valid ? __fdiv_rn(st[P.R * D + r * C + c] - c_mean[c], c_std[c]) : 0.f;
        for (int l = 0; l < L; ++l) sm[P.o_xin + l * P.xin_stride + (D + c) * RP + r] = v;
      }
      for (int idx = tid; idx < P.DP * RP; idx += NT) ld[idx] = 0.f;
    }
    __syncthreads();
    if (ti + 1 < my_n) stage_issue(ti + 1);

    // ---- forward through the L MAF layers
    for (int k = 0; k < L; ++k) {
      const float* wf = w_acquire(req++);
      float* slot = sm + P.o_slot + ((BWD && P.stash_all) ? k : 0) * P.slot_stride;
      layer_forward(k, wf, slot, false);
    }

    // ---- log-probability per row (jaxili/model.py:919-921 + 1094-1098); seed of the reverse pass
    for (int r = tid; r < Rt; r += NT) {
      float acc = 0.f;
      for (int j = 0; j < D; ++j) {
        const float u = gu[j * RP + r];
        acc += ld[j * RP + r] - 0.5f * u * u;
        if (BWD) gu[j * RP + r] = r < valid ? u * A.inv_b : 0.f;
      }
      const float lp = acc - (float)D * kHalfLog2Pi + A.logdet_const;
      if (BWD) {
        if (r < valid) atomicAdd(loss_acc, -lp * A.inv_b);
      } else if (r < valid) {
        A.out_logp[row0 + r] = lp;
      }
    }
    __syncthreads();

    if (BWD) {
      float* pg_cta = A.partial_grad + (size_t)blockIdx.x * P.P;
      const bool accum = ti > 0;
      for (int k = L - 1; k >= 0; --k) {
        const float* wimg_s = w_acquire(req++);
        const float* wb = recompute ? wimg_s + P.fwd_floats : wimg_s;
        float* slot = sm + P.o_slot + (P.stash_all ? k : 0) * P.slot_stride;
        if (recompute) layer_forward(k, wimg_s, slot, true);
        float* pg = pg_cta + (size_t)k * P.ppl;
        const float* xin = sm + P.o_xin + k * P.xin_stride;

        // gradient w.r.t. the MADE outputs (mu | logp) and the direct path to x
        {
          const float* sbuf = slot + P.s_off;
          const float* vbuf = slot + P.v_off;
          for (int idx = tid; idx < D * Rt; idx += NT) {
            const int j = idx / Rt, r = idx - j * Rt;
            const int jj = P.reverse ? D - 1 - j : j;
            const float gv = gu[jj * RP + r];
            const float s = sbuf[j * RP + r], v = vbuf[j * RP + r];
            const bool ok = r < valid;
            obuf[j * RP + r] = ok ? -gv * s : 0.f;
            obuf[(D + j) * RP + r] = ok ? 0.5f * gv * v - 0.5f * A.inv_b : 0.f;
            gx[j * RP + r] = ok ? gv * s : 0.f;
          }
        }
        __syncthreads();

        for (int i = n_lin - 1; i >= 0; --i) {
          const LinDesc& ln = P.lin[i];
          const float* G = (i == n_lin - 1) ? obuf : sm + P.o_g[i & 1];
          const float* Ain = (i == 0) ? xin : slot + P.h_off[i];
          const float* Wt = wb + ln.wt_off;
          // three warp-granular task lists share the phase: X (input gradient), B (weight
          // gradient) and the bias gradient (one warp-shuffle reduction per output feature)
          const int xcols = (i == 0) ? P.DP : ln.KP;
          const int nXR = Rt / XM;
          const int nX = nXR * (xcols / XN);
          const int nKB = (ln.K + BK - 1) / BK, nNB = (ln.N + BN - 1) / BN;
          const int nB = nKB * nNB;
          const int bX = (nX + 31) >> 5, bB = (nB + 31) >> 5;
          const int total = bX + bB + ln.N;
          for (int blk = warp; blk < total; blk += NW) {
            if (blk < bB) {
              const int it = blk * 32 + lane;
              if (it < nB)
                mm_inner<BK, BN>(Ain, ln.K, G, ln.N, Rt, RP, it / nNB, nKB, it % nNB, nNB,
                                 pg + ln.pw_off, accum);
            } else if (blk < bB + bX) {
              const int it = (blk - bB) * 32 + lane;
              if (it < nX) {
                const int r0 = (it % nXR) * XM, c0 = (it / nXR) * XN;
                if (i > 0) {
                  float* gout = sm + P.o_g[(i - 1) & 1];
                  const float* hprev = slot + P.h_off[i];
                  const float* dprev = slot + P.d_off[i];
                  mm_outer<XM, XN>(G, ln.N, RP, Wt, ln.KP, r0, c0, [&](float (&acc)[XN][XM]) {
#pragma unroll
                    for (int j = 0; j < XN; ++j) {
                      float dv[XM], gv[XM];
                      ldv(dv, (RELU ? hprev : dprev) + (c0 + j) * RP + r0);
#pragma unroll
                      for (int q = 0; q < XM; ++q)
                        gv[q] = RELU ? (dv[q] > 0.f ? acc[j][q] : 0.f) : acc[j][q] * dv[q];
                      stv(gout + (c0 + j) * RP + r0, gv);
                    }
                  });
                } else {
                  mm_outer<XM, XN>(G, ln.N, RP, Wt, ln.KP, r0, c0, [&](float (&acc)[XN][XM]) {
#pragma unroll
                    for (int j = 0; j < XN; ++j) {
                      float dv[XM], gv[XM];
                      ldv(dv, gx + (c0 + j) * RP + r0);
#pragma unroll
                      for (int q = 0; q < XM; ++q) gv[q] = acc[j][q] + dv[q];
                      stv(gu + (c0 + j) * RP + r0, gv);
                    }
                  });
                }
              }
            } else {
              const int n = blk - bB - bX;
              float sacc = 0.f;
              for (int r = lane; r < Rt; r += 32) sacc += G[n * RP + r];
              sacc = warp_sum(sacc);
              if (lane == 0) {
                float* p = pg + ln.pb_off + n;
                *p = accum ? *p + sacc : sacc;
              }
            }
          }
          __syncthreads();
        }
      }
    }
  }
  if (BWD && tid == 0) A.partial_loss[blockIdx.x] = *loss_acc;
}

}  // namespace mafb
