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
  float* out_logp;            // forward-only: [B] (nullable)
  float* out_u;               // forward-only: [B][D] base-space point u (nullable)
  float* out_logdet;          // forward-only: [B] summed log|det J| of the MAF (nullable)
  float* out_dy;              // reverse pass: [B][C] gradient w.r.t. the RAW conditioner rows (nullable);
                              // with it the input-gradient product of the first linear covers the
                              // conditioner columns and no weight gradients are scheduled
  float* partial_grad;        // training: [grid][P]
  float* partial_loss;        // training: [grid]
  int Rt;                     // rows per tile in this launch (multiple of 4, <= Plan::R)
  int rb_shift;               // log2 of the power of two >= Rt/4 (row blocks per warp slot group)
  int n_tiles;
  // per-linear work-list geometry, precomputed on the host (no integer division in the kernel)
  struct Phase {
    int fB;                   // forward: warp blocks
    int xC, xB;               // input gradient: column blocks, warp blocks
    int nKB, nNB;             // weight gradient: tile grid ((K+1)/4 x N/4, rounded up)
    int nb_shift, nbg, nKG;   // column blocks per half-warp = 1<<nb_shift (<= 8), groups, row groups
    int bB;                   // weight gradient: warp blocks = nKG * nbg
  } ph[MAFB_MAX_LIN];
  // chain kernel (chain_kernel.cuh): per product of the row-owning warp pairs
  struct ChainPh {
    int nsteps;               // reduction steps per lane = padded reduction length / 4
    int nCB;                  // column blocks of the product
    int split;                // 1: the pair's warps split the column blocks; 0: narrow product, one warp does it
  } cf[MAFB_MAX_LIN], cx[MAFB_MAX_LIN];   // forward product, input-gradient product
  long long* trace;           // debug: clock64 stamps of CTA 0 at phase boundaries (nullable)
};

// Every micro-GEMM uses 4x4 register tiles.  The reduction dimension is split over the two
// half-warps and recombined with a warp-shuffle reduce-scatter: twice the parallel work units and
// half the dependent chain of one tile per thread, which the small per-SM row count (28 rows at
// B=4096) needs.  The lane -> tile maps are chosen for shared-memory MULTICAST: lanes of one load
// instruction share operand rows (8 row blocks x 2 column blocks for h.W and G.W^T; 2 x 8 tile
// groups for [A;1]^T G), which keeps the kernel at >= 5 FFMA per LSU wavefront -- the 128 B/clk
// shared-memory port, not the FMA pipe, is what a naive mapping saturates first.  Work units are
// mapped to (warp, lane) with shifts and masks only; the work-list geometry comes from the host.
constexpr int TS = 4;           // tile side
constexpr int KS = 2;           // lane groups splitting the reduction of outer-form products
constexpr float kHalfLog2Pi = 0.9189385332046727f;

// NLIN > 0 fixes the number of masked linears per MADE at compile time (loops over them unroll and
// the per-linear descriptors become constant-bank operands); NLIN == 0 reads it from the plan.
template <bool BWD, bool RELU, int NT, int NLIN>
__global__ void __launch_bounds__(NT, NT <= 256 ? 2 : 1)
maf_flow_kernel(const __grid_constant__ Plan P, const __grid_constant__ FlowArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem_raw);   // [2] weight images
  uint64_t* sbar = wbar + 2;                                // [2] input tiles
  float* loss_acc = reinterpret_cast<float*>(smem_raw + 32);
  float* sm = reinterpret_cast<float*>(smem_raw + MAFB_SMEM_HDR);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const int D = P.D, C = P.C, L = P.L, RP = P.RP;
  const int n_lin = NLIN > 0 ? NLIN : P.n_lin;
  const int Rt = A.Rt;
  const int nRB = Rt / TS;
  // outer-form products: a warp holds 16 tile slots x 2 reduction lanes; slot -> (row block, col block)
  const int kq = lane >> 4;
  const int slot16 = lane & 15;
  const int o_rb = slot16 & ((1 << A.rb_shift) - 1);
  const int o_cbl = slot16 >> A.rb_shift;         // column block within the warp's group
  const int o_cpw = 16 >> A.rb_shift;             // column blocks per warp
  const bool o_rok = o_rb < nRB;
  const int o_r0 = (o_rok ? o_rb : 0) * TS;
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

#ifdef MAFB_TRACE
  int n_trace = 0;
  auto stamp = [&](int id) {
    if (A.trace && blockIdx.x == 0 && tid == 0 && n_trace < 500) {
      A.trace[2 * n_trace] = id;
      A.trace[2 * n_trace + 1] = clock64();
      ++n_trace;
    }
  };
#else
  auto stamp = [](int) {};
#endif
  stamp(0);

  const bool recompute = BWD && !P.stash_all;
  const bool two_wbuf = P.n_wbuf == 2;
  const int per_tile = BWD ? 2 * L : L;
  const int total_req = my_n * per_tile;

  // ---- weight image requests: s-th request of this CTA.  Issued by the last warp's lane 0 (that
  //      warp is the least loaded in every phase) so the issue latency stays off the critical path.
  auto w_issue = [&](int s) {
    const int j = s % per_tile;
    int layer, part;
    if (!BWD || j < L) { layer = j; part = 0; } else { layer = 2 * L - 1 - j; part = 1; }
    const float* src = A.wimg + (size_t)layer * P.img_floats;
    uint32_t bytes;
    if (part == 0) bytes = P.fwd_floats * 4u;
    else if (recompute) bytes = P.img_floats * 4u;
    else { src += P.fwd_floats; bytes = P.bwd_floats * 4u; }
    const int buf = two_wbuf ? (s & 1) : 0;
    mbar_expect_tx(&wbar[buf], bytes);
    bulk_g2s(sm + P.o_wbuf + buf * P.wbuf_stride, src, bytes, &wbar[buf]);
  };
  const bool issuer = tid == NT - 32;
  // every thread: make request s resident, prefetch s+1; returns the buffer.  The buffer being
  // refilled was last read before the __syncthreads that precedes this call.
  auto w_acquire = [&](int s) -> const float* {
    const int buf = two_wbuf ? (s & 1) : 0;       // n_wbuf is 1 or 2: no integer division on this path
    if (issuer) {
      if (P.n_wbuf == 1) w_issue(s);
      else if (s + 1 < total_req) w_issue(s + 1);
    }
    mbar_wait(&wbar[buf], (uint32_t)(two_wbuf ? (s >> 1) & 1 : s & 1));
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
      if (issuer) {
        mbar_expect_tx(&sbar[i & 1], (uint32_t)(valid * (D + C) * 4));
        bulk_g2s(dst, xs, (uint32_t)(valid * D * 4), &sbar[i & 1]);
        if (C) bulk_g2s(dst + P.R * D, ys, (uint32_t)(valid * C * 4), &sbar[i & 1]);
      }
    } else {
      for (int idx = tid; idx < valid * D; idx += NT) dst[idx] = xs[idx];
      for (int idx = tid; idx < valid * C; idx += NT) dst[P.R * D + idx] = ys[idx];
      if (issuer) mbar_arrive(&sbar[i & 1]);
    }
  };

  if (issuer && P.n_wbuf == 2 && total_req > 0) w_issue(0);
  if (my_n > 0) stage_issue(0);

  float* const obuf = sm + P.o_obuf;
  float* const gu = sm + P.o_gu;
  float* const gx = sm + P.o_gx;
  float* const ld = sm + P.o_ld;
  float* const cst = sm + P.o_cst;   // [shift D][inv_scale D][mean C][std C]
  for (int idx = tid; idx < 2 * D + 2 * C; idx += NT) cst[idx] = A.stdc[idx];

  if (BWD) {
    // a feature row of ones behind every weight-gradient A operand: [A;1]^T G yields the bias
    // gradient as one more row of the weight-gradient product.  Written once, never overwritten.
    const int n_slots = P.stash_all ? L : 1;
    for (int idx = tid; idx < RP; idx += NT) {
      for (int l = 0; l < L; ++l) sm[P.o_xin + l * P.xin_stride + P.lin[0].KP * RP + idx] = 1.f;
      for (int sl = 0; sl < n_slots; ++sl)
        for (int i = 1; i < n_lin; ++i)
          sm[P.o_slot + sl * P.slot_stride + P.h_off[i] + P.lin[i].KP * RP + idx] = 1.f;
    }
  }

  // ---- one MADE + affine transform.  replay: recomputation inside the reverse pass
  //      (no log-det accumulation, no hand-off of v to the next layer).
  auto layer_forward = [&](int k, const float* wf, float* slot, bool replay) {
    const float* in = sm + P.o_xin + k * P.xin_stride;
    constexpr int kUnrollLin = NLIN > 0 ? NLIN : 1;     // NLIN == 0: rolled loop over the linears
#pragma unroll kUnrollLin
    for (int i = 0; i < (NLIN > 0 ? NLIN : n_lin); ++i) {
      const LinDesc& ln = P.lin[i];
      const float* W = wf + ln.w_off;
      const float* bias = wf + ln.b_off;
      const int nCB = ln.NP / TS;
      const int Q = ln.K, CP = ln.NP;
      const bool last = i == n_lin - 1;
      float* hout = last ? obuf : slot + P.h_off[i + 1];
      float* dout = slot + P.d_off[last ? 1 : i + 1];
      const float* xin_k = sm + P.o_xin + k * P.xin_stride;
      float* xnext = (k + 1 < L) ? sm + P.o_xin + (k + 1) * P.xin_stride : gu;
      float* sbuf = slot + P.s_off;
      float* vbuf = slot + P.v_off;
      for (int blk = warp; blk < A.ph[i].fB; blk += NW) {
        const int cb = blk * o_cpw + o_cbl;
        const bool ok = o_rok && cb < nCB;
        const int c0 = (cb < nCB ? cb : nCB - 1) * TS;
        float acc[TS][TS], res[2][TS];
        stamp(30);
        mm_outer_acc<TS, TS>(acc, in, Q, RP, W, CP, o_r0, c0, kq, KS);
        stamp(31);
        reduce_scatter2(acc, kq, res);
        stamp(32);
        if (ok) {   // half-warp kq finishes tile columns 2kq, 2kq+1
          if (!last) {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int c = c0 + 2 * kq + cc;
              const float bj = bias[c];
              float hv[TS], dv[TS];
#pragma unroll
              for (int q = 0; q < TS; ++q) act_eval<RELU>(P.act, res[cc][q] + bj, hv[q], dv[q]);
              stv(hout + c * RP + o_r0, hv);
              if (!RELU && BWD) stv(dout + c * RP + o_r0, dv);
            }
          } else {
            // the last linear's image interleaves (mu_j, logp_j): this thread holds both for feature j and
            // applies the affine autoregressive transform (jaxili/model.py:739-742) right here
            const int c = c0 + 2 * kq;
            const int j = c >> 1;
            if (j < D) {
              const float bm = bias[c], bl = bias[c + 1];
              float xv[TS], sv[TS], vv[TS];
              ldv(xv, xin_k + j * RP + o_r0);
#pragma unroll
              for (int q = 0; q < TS; ++q) {
                const float lp = res[1][q] + bl;
                sv[q] = expf(0.5f * lp);
                vv[q] = (xv[q] - (res[0][q] + bm)) * sv[q];
                res[1][q] = 0.5f * lp;
              }
              if (BWD) {
                stv(sbuf + j * RP + o_r0, sv);
                stv(vbuf + j * RP + o_r0, vv);
              }
              if (!replay) {
                float lv[TS];
                ldv(lv, ld + j * RP + o_r0);
#pragma unroll
                for (int q = 0; q < TS; ++q) lv[q] += res[1][q];
                stv(ld + j * RP + o_r0, lv);
                const int jj = P.reverse ? D - 1 - j : j;
                stv(xnext + jj * RP + o_r0, vv);
              }
            }
          }
        }
      }
      in = hout;
      stamp(33);
      __syncthreads();
      stamp(10 + i);
    }
  };

  for (int ti = 0; ti < my_n; ++ti) {
    const int t = blockIdx.x + ti * grid;
    const long long row0 = (long long)t * Rt;
    const int valid = (int)min((long long)Rt, A.B - row0);
    int req = ti * per_tile;

    // ---- stage in: wait for the bulk copy, transpose to feature-major, standardise
    mbar_wait(&sbar[ti & 1], (uint32_t)((ti >> 1) & 1));
    __syncthreads();
    stamp(1);
    {
      const float* st = sm + P.o_stage + (ti & 1) * P.stage_stride;
      float* xin0 = sm + P.o_xin;
      for (int r = lane; r < Rt; r += 32) {
        const bool ok = r < valid;
        for (int j = warp; j < D + C; j += NW) {
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
    }
    if (BWD && A.out_dy)
      for (int idx = tid; idx < (P.lin[0].KP - P.DP + 4) * RP; idx += NT) sm[P.o_gy + idx] = 0.f;
    __syncthreads();
    stamp(2);
    if (ti + 1 < my_n) stage_issue(ti + 1);

    // ---- forward through the L MAF layers
    for (int k = 0; k < L; ++k) {
      const float* wf = w_acquire(req++);
      stamp(3);
      float* slot = sm + P.o_slot + ((BWD && P.stash_all) ? k : 0) * P.slot_stride;
      layer_forward(k, wf, slot, false);
    }

    // ---- log-probability per row (jaxili/model.py:919-921 + 1094-1098); seed of the reverse pass
    if (BWD) {
      // all warps: RPW rows per warp, LPR lanes per row striding the features; the row sums go through an xor
      // shuffle tree (fixed order).  With every layer's (s, v) stashed the last layer's elementwise reverse step
      // (gradient w.r.t. mu | logp and the direct path) is done right here, saving a phase.
      constexpr int RPW = 32 / NW, LPR = 32 / RPW;        // 512 threads: 2 x 16, 256 threads: 4 x 8
      const int r = warp * RPW + lane / LPR;
      const int f0 = lane % LPR;
      const bool rok = r < Rt, vok = r < valid;
      const bool fuse_fin = P.stash_all != 0;
      const float* pslot = sm + P.o_slot + (P.stash_all ? L - 1 : 0) * P.slot_stride;
      float* gx_l = gx + ((L - 1) & 1) * P.DP * RP;
      float acc = 0.f;
      if (rok)
        for (int j = f0; j < D; j += LPR) {
          const float u = gu[j * RP + r];
          const float l = ld[j * RP + r];
          acc += l - 0.5f * u * u;
          const float gv = vok ? u * A.inv_b : 0.f;
          gu[j * RP + r] = gv;
          if (fuse_fin) {
            const int jm = P.reverse ? D - 1 - j : j;      // output j of the flow is MADE feature jm of the last layer
            const float s = pslot[P.s_off + jm * RP + r], v = pslot[P.v_off + jm * RP + r];
            obuf[jm * RP + r] = vok ? -gv * s : 0.f;
            obuf[(D + jm) * RP + r] = vok ? 0.5f * gv * v - 0.5f * A.inv_b : 0.f;
            gx_l[jm * RP + r] = vok ? gv * s : 0.f;
          }
        }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const float lp = acc - (float)D * kHalfLog2Pi + A.logdet_const;
      const bool lead = f0 == 0 && vok;
      if (lead && A.out_logp) A.out_logp[row0 + r] = lp;
      const float contrib = warp_sum(lead ? -lp * A.inv_b : 0.f);
      if (lane == 0) sm[P.o_red + warp] = contrib;
    }
    if (!BWD) {
      for (int r = tid; r < Rt; r += NT) {   // forward only: lane = row
        float acc = 0.f, ldsum = 0.f;
        for (int j = 0; j < D; ++j) {
          const float u = gu[j * RP + r];
          const float l = ld[j * RP + r];
          ldsum += l;
          acc += l - 0.5f * u * u;
        }
        const float lp = acc - (float)D * kHalfLog2Pi + A.logdet_const;
        if (r < valid) {
          if (A.out_logp) A.out_logp[row0 + r] = lp;
          if (A.out_logdet) A.out_logdet[row0 + r] = ldsum;
        }
      }
    }
    if (!BWD && A.out_u) {   // u rows are contiguous in HBM: coalesced store
      float* dst = A.out_u + row0 * D;
      for (int idx = tid; idx < valid * D; idx += NT) {
        const int r = idx / D, j = idx - r * D;
        dst[idx] = gu[j * RP + r];
      }
    }
    __syncthreads();
    if (BWD && tid == 0) {      // per-warp loss contributions in warp order: reproducible
      float t = 0.f;
      for (int w = 0; w < NW; ++w) t += sm[P.o_red + w];
      *loss_acc += t;
    }
    stamp(6);

    if (BWD) {
      float* pg_cta = A.partial_grad + (size_t)blockIdx.x * P.pgs;
      const bool accum = ti > 0;
      for (int k = L - 1; k >= 0; --k) {
        const float* wimg_s = w_acquire(req++);
        stamp(4);
        const float* wb = recompute ? wimg_s + P.fwd_floats : wimg_s;
        float* slot = sm + P.o_slot + (P.stash_all ? k : 0) * P.slot_stride;
        if (recompute) layer_forward(k, wimg_s, slot, true);
        float* pg = pg_cta + (size_t)k * P.ppl;
        const float* xin = sm + P.o_xin + k * P.xin_stride;

        // gradient w.r.t. the MADE outputs (mu | logp) and the direct path to x.  With every layer's (s, v)
        // stashed this never runs as its own phase: the last layer's is part of the finalize step above, the
        // others are fused into the epilogue of the following layer's input-gradient product (below).
        const bool fuse_ew = P.stash_all != 0;
        float* gx_k = gx + (k & 1) * P.DP * RP;
        if (!fuse_ew) {
          const float* sbuf = slot + P.s_off;
          const float* vbuf = slot + P.v_off;
          for (int r = lane; r < Rt; r += 32) {
            const bool ok = r < valid;
            for (int j = warp; j < D; j += NW) {
              const int jj = P.reverse ? D - 1 - j : j;
              const float gv = gu[jj * RP + r];
              const float s = sbuf[j * RP + r], v = vbuf[j * RP + r];
              obuf[j * RP + r] = ok ? -gv * s : 0.f;
              obuf[(D + j) * RP + r] = ok ? 0.5f * gv * v - 0.5f * A.inv_b : 0.f;
              gx_k[j * RP + r] = ok ? gv * s : 0.f;
            }
          }
          __syncthreads();
          stamp(5);
        }

        constexpr int kUnrollLinB = NLIN > 0 ? NLIN : 1;
#pragma unroll kUnrollLinB
        for (int ii = 0; ii < (NLIN > 0 ? NLIN : n_lin); ++ii) {
          const int i = n_lin - 1 - ii;
          const LinDesc& ln = P.lin[i];
          const float* G = (i == n_lin - 1) ? obuf : sm + P.o_g[i & 1];
          const float* Ain = (i == 0) ? xin : slot + P.h_off[i];
          const float* Wt = wb + ln.wt_off;
          // two warp-granular work lists share the phase:
          //   X: input gradient g_in = G (M*W)^T; a warp holds 16 tiles, its half-warps split the sum over n
          //   B: weight + bias gradient [A;1]^T G; a warp holds (16 >> nb_shift) x (1 << nb_shift) tiles,
          //      its half-warps split the batch rows
          const FlowArgs::Phase& ph = A.ph[i];
          const int N = ln.N, K = ln.K, KP = ln.KP;
          const int bX = ph.xB, nXC = ph.xC;
          for (int blk = warp; blk < bX + ph.bB; blk += NW) {
            stamp(blk < bX ? 40 : 42);
            if (blk < bX) {
              const int cb = blk * o_cpw + o_cbl;
              const bool ok = o_rok && cb < nXC;
              const int c0 = (cb < nXC ? cb : nXC - 1) * TS;
              float acc[TS][TS], res[2][TS];
              mm_outer_acc<TS, TS>(acc, G, N, RP, Wt, KP, o_r0, c0, kq, KS);
              reduce_scatter2(acc, kq, res);
              if (ok) {
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                  const int c = c0 + 2 * kq + cc;
                  float dv[TS], gv[TS];
                  if (i > 0) {   // through the activation: g_a = g_h * act'(a)
                    ldv(dv, slot + (RELU ? P.h_off[i] : P.d_off[i]) + c * RP + o_r0);
#pragma unroll
                    for (int q = 0; q < TS; ++q)
                      gv[q] = RELU ? (dv[q] > 0.f ? res[cc][q] : 0.f) : res[cc][q] * dv[q];
                    stv(sm + P.o_g[(i - 1) & 1] + c * RP + o_r0, gv);
                  } else if (c < D || !A.out_dy) {   // gradient w.r.t. this layer's input x: MADE + direct path
                    ldv(dv, gx_k + c * RP + o_r0);
#pragma unroll
                    for (int q = 0; q < TS; ++q) gv[q] = res[cc][q] + dv[q];
                    if (!fuse_ew) {
                      stv(gu + c * RP + o_r0, gv);
                    } else if (k > 0 && c < D) {
                      // x_k is layer k-1's (flipped) output: continue straight into that layer's elementwise step
                      const int j = P.reverse ? D - 1 - c : c;
                      const float* pslot = sm + P.o_slot + (k - 1) * P.slot_stride;
                      float sv[TS], vv[TS], o0[TS], o1[TS], gn[TS];
                      ldv(sv, pslot + P.s_off + j * RP + o_r0);
                      ldv(vv, pslot + P.v_off + j * RP + o_r0);
#pragma unroll
                      for (int q = 0; q < TS; ++q) {
                        const bool okr = o_r0 + q < valid;
                        o0[q] = okr ? -gv[q] * sv[q] : 0.f;
                        o1[q] = okr ? 0.5f * gv[q] * vv[q] - 0.5f * A.inv_b : 0.f;
                        gn[q] = okr ? gv[q] * sv[q] : 0.f;
                      }
                      stv(obuf + j * RP + o_r0, o0);
                      stv(obuf + (D + j) * RP + o_r0, o1);
                      stv(gx + ((k - 1) & 1) * P.DP * RP + j * RP + o_r0, gn);
                    }
                  } else if (c < P.K0) {             // dy mode, conditioner columns: summed over the layers
                    float* gyp = sm + P.o_gy + (c - D) * RP + o_r0;
                    ldv(dv, gyp);
#pragma unroll
                    for (int q = 0; q < TS; ++q) gv[q] = res[cc][q] + dv[q];
                    stv(gyp, gv);
                  }
                }
              }
            } else {
              int b = blk - bX, g = 0;
              while (b >= ph.nKG) { b -= ph.nKG; ++g; }        // g < nbg <= a few
              const int l16 = lane & 15;
              const int kb = b * (16 >> ph.nb_shift) + (l16 >> ph.nb_shift);
              const int nb = (g << ph.nb_shift) + (l16 & ((1 << ph.nb_shift) - 1));
              mm_inner_rs2(Ain, K, KP, G, N, Rt, RP, kb, ph.nKB, nb < ph.nNB ? nb : ph.nNB - 1, ph.nNB, lane >> 4,
                           nb < ph.nNB && kb < ph.nKB, pg + ln.pw_off, accum);
            }
            stamp(blk < bX ? 41 : 43);
          }
          __syncthreads();
          stamp(20 + i);
        }
      }
      if (A.out_dy) {   // d/dy = (d/dy') / std (Standardizer), rows are contiguous in HBM: coalesced store
        float* dst = A.out_dy + row0 * C;
        for (int idx = tid; idx < valid * C; idx += NT) {
          const int r = idx / C, c = idx - r * C;
          dst[idx] = __fdiv_rn(sm[P.o_gy + c * RP + r], cst[2 * D + C + c]);
        }
        __syncthreads();
      }
    }
  }
  stamp(99);
#ifdef MAFB_TRACE
  if (A.trace && blockIdx.x == 0 && tid == 0) A.trace[1000] = n_trace;
#endif
  if (BWD && tid == 0) A.partial_loss[blockIdx.x] = *loss_acc;
}

}  // namespace mafb
