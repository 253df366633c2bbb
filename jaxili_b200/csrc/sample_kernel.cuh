// sample_kernel.cuh -- autoregressive inversion (posterior sampling) for sm_100a.
//
// Reference semantics (jaxili/model.py:745-773, 876-901, 923-947, 1116-1126): for every MAF layer,
// last to first, D sequential full MADE passes; pass d fixes x[:,d] = mu_d + exp(min(-logp_d/2, 10)) u_d.
//
// Here each warp owns 32 sample rows for the whole flow and never leaves shared memory.  Because
// the masks come from degrees (jaxili/model.py:603-663), a hidden unit of degree m is final as soon
// as x_0..x_m are known.  Hidden units are stored SORTED BY DEGREE (done by the packer), so pass d
// only evaluates the units of degree d-1 -- a contiguous column group whose fan-in is a contiguous
// prefix of the previous level -- and the two outputs (mu_d, logp_d).  All D passes of a layer
// together cost about one masked MADE evaluation instead of D dense ones; the values are the same
// sums of the same non-zero terms.  (Disclosed in DESIGN.md: algorithmic FLOPs are quoted for the
// reference's dense D-pass formulation.)
#pragma once
#include <curand_kernel.h>
#include "common.cuh"

namespace mafb {

struct SampleArgs {
  const float* simg;               // [L][img_floats] degree-sorted images
  const float* u;                  // [N][D] base noise, or nullptr -> Philox
  unsigned long long seed, row_offset;
  const float* y_obs;              // [n_obs][C]
  long long n_obs, rows_per_obs, N;
  const float* stdc;               // [shift D][inv_scale D][mean C][std C][scale D]
  float* out;                      // [N][D]
  float* out_logdet;               // [N] sum over layers of sum_d min(-logp_d/2, 10) (nullable)
};

constexpr int SM_ = 4, SN_ = 2;    // sampler tile: 4 rows x 2 columns per lane; lanes = 8 row blocks x 4 column blocks

template <bool RELU>
__global__ void __launch_bounds__(256, 1)
maf_sample_kernel(const __grid_constant__ SamplePlan S, const __grid_constant__ SampleArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* ibar = reinterpret_cast<uint64_t*>(smem_raw);
  float* sm = reinterpret_cast<float*>(smem_raw + 64);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NWS = blockDim.x >> 5;
  const int D = S.D, C = S.C, L = S.L, RP = S.RP, n_hidden = S.n_lin - 1;
  const int r0 = (lane & 7) * SM_, cb = lane >> 3;
  const long long total_tiles = (A.N + 31) / 32;
  const long long per_iter = (long long)gridDim.x * NWS;
  const int n_iter = (int)((total_tiles + per_iter - 1) / per_iter);
  const int total_req = n_iter * L;

  if (tid == 0) {
    mbar_init(&ibar[0], 1);
    mbar_init(&ibar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto img_issue = [&](int s) {   // thread 0
    const int k = L - 1 - (s % L);
    const int buf = s % S.n_img;
    fence_proxy_async();
    mbar_expect_tx(&ibar[buf], S.img_floats * 4u);
    bulk_g2s(sm + S.o_img + buf * S.img_floats, A.simg + (size_t)k * S.img_floats, S.img_floats * 4u,
             &ibar[buf]);
  };
  if (tid == 0 && S.n_img == 2 && total_req > 0) img_issue(0);

  float* wbase = sm + S.o_warp + warp * S.warp_stride;
  float* ybuf = wbase + S.y_off;
  const float* c_shift = A.stdc;
  const float* c_mean = A.stdc + 2 * D;
  const float* c_std = A.stdc + 2 * D + C;
  const float* c_scale = A.stdc + 2 * D + 2 * C;

  for (int it = 0; it < n_iter; ++it) {
    const long long tile = ((long long)it * gridDim.x + blockIdx.x) * NWS + warp;
    const bool active = tile < total_tiles;
    const long long row0 = tile * 32;
    const int valid = active ? (int)min(32LL, A.N - row0) : 0;
    float* ubuf = wbase + S.u_off;
    float* xbuf = wbase + S.x_off;
    float ldinv[SM_];
#pragma unroll
    for (int i = 0; i < SM_; ++i) ldinv[i] = 0.f;

    if (active) {
      // base noise -> feature-major
      if (A.u) {
        const float* src = A.u + row0 * D;
        for (int idx = lane; idx < 32 * D; idx += 32) {
          const int r = idx / D, j = idx - r * D;
          ubuf[j * RP + r] = r < valid ? src[idx] : 0.f;
        }
      } else {
        curandStatePhilox4_32_10_t st;
        curand_init(A.seed, A.row_offset + (unsigned long long)(row0 + lane), 0ULL, &st);
        for (int j = 0; j < D; j += 4) {
          const float4 z = curand_normal4(&st);
          ubuf[j * RP + lane] = z.x;
          if (j + 1 < D) ubuf[(j + 1) * RP + lane] = z.y;
          if (j + 2 < D) ubuf[(j + 2) * RP + lane] = z.z;
          if (j + 3 < D) ubuf[(j + 3) * RP + lane] = z.w;
        }
      }
      // conditioning rows: Standardizer (y - mean) / std of the row's observation
      {
        const long long row = row0 + min(lane, max(valid - 1, 0));
        long long obs = row / A.rows_per_obs;
        if (obs >= A.n_obs) obs = A.n_obs - 1;
        for (int c = 0; c < C; ++c)
          ybuf[c * RP + lane] = __fdiv_rn(A.y_obs[obs * C + c] - c_mean[c], c_std[c]);
      }
    }
    __syncwarp();

    for (int li = 0; li < L; ++li) {
      const int s = it * L + li;
      const int buf = s % S.n_img;
      if (tid == 0) {
        if (S.n_img == 1) img_issue(s);
        else if (s + 1 < total_req) img_issue(s + 1);
      }
      mbar_wait(&ibar[buf], (uint32_t)((s / S.n_img) & 1));
      const float* img = sm + S.o_img + buf * S.img_floats;

      if (active) {
        const int* T = reinterpret_cast<const int*>(img + S.tab_off);
        float* lv1 = wbase + S.lvl_off[1];
        const float* W0 = img + S.w_off[0];
        const float* b0 = img + S.b_off[0];
        const int w1 = S.width[1];
        // level-1 pre-activations from the conditioning inputs and the bias (x-independent)
        for (int c0 = cb * SN_; c0 < w1; c0 += 4 * SN_) {
          mm_outer<SM_, SN_>(ybuf, C, RP, W0 + D * w1, w1, r0, c0, [&](float (&acc)[SN_][SM_]) {
#pragma unroll
            for (int j = 0; j < SN_; ++j) {
              const float bj = b0[c0 + j];
              float av[SM_];
#pragma unroll
              for (int q = 0; q < SM_; ++q) av[q] = acc[j][q] + bj;
              stv(lv1 + (c0 + j) * RP + r0, av);
            }
          });
        }
        __syncwarp();

        for (int d = 0; d < D; ++d) {
          if (d > 0) {
            const int m = d - 1;
            {  // level 1, units of degree m: fan-in x_0..x_m
              const int g0 = T[m], g1 = T[m + 1];
              for (int c0 = g0 + cb * SN_; c0 < g1; c0 += 4 * SN_) {
                mm_outer<SM_, SN_>(xbuf, d, RP, W0, w1, r0, c0, [&](float (&acc)[SN_][SM_]) {
#pragma unroll
                  for (int j = 0; j < SN_; ++j) {
                    float av[SM_], hv[SM_], dv;
                    ldv(av, lv1 + (c0 + j) * RP + r0);
#pragma unroll
                    for (int q = 0; q < SM_; ++q) act_eval<RELU>(S.act, acc[j][q] + av[q], hv[q], dv);
                    stv(lv1 + (c0 + j) * RP + r0, hv);
                  }
                });
              }
              __syncwarp();
            }
            for (int l = 2; l <= n_hidden; ++l) {  // deeper levels: fan-in = sorted prefix of level l-1
              const int* Tl = T + (l - 1) * (D + 1);
              const int g0 = Tl[m], g1 = Tl[m + 1];
              const int Q = T[(l - 2) * (D + 1) + m + 1];
              const float* W = img + S.w_off[l - 1];
              const float* b = img + S.b_off[l - 1];
              const int wl = S.width[l];
              const float* hin = wbase + S.lvl_off[l - 1];
              float* hout = wbase + S.lvl_off[l];
              for (int c0 = g0 + cb * SN_; c0 < g1; c0 += 4 * SN_) {
                mm_outer<SM_, SN_>(hin, Q, RP, W, wl, r0, c0, [&](float (&acc)[SN_][SM_]) {
#pragma unroll
                  for (int j = 0; j < SN_; ++j) {
                    const float bj = b[c0 + j];
                    float hv[SM_], dv;
#pragma unroll
                    for (int q = 0; q < SM_; ++q) act_eval<RELU>(S.act, acc[j][q] + bj, hv[q], dv);
                    stv(hout + (c0 + j) * RP + r0, hv);
                  }
                });
              }
              __syncwarp();
            }
          }
          // outputs (mu_d, logp_d): fan-in = last-level units of degree < d; the four column
          // blocks split the reduction and combine with warp shuffles
          {
            const int Q = d > 0 ? T[(n_hidden - 1) * (D + 1) + d] : 0;
            const float* Wo = img + S.w_off[n_hidden] + 2 * d;
            const float* bo = img + S.b_off[n_hidden] + 2 * d;
            const int wo = S.width[n_hidden + 1];
            const float* hin = wbase + S.lvl_off[n_hidden];
            float acc[2][SM_];
#pragma unroll
            for (int q = 0; q < SM_; ++q) acc[0][q] = acc[1][q] = 0.f;
#pragma unroll 2
            for (int q = cb; q < Q; q += 4) {
              float hv[SM_], wv[2];
              ldv(hv, hin + q * RP + r0);
              ldv(wv, Wo + q * wo);
#pragma unroll
              for (int i = 0; i < SM_; ++i) {
                acc[0][i] = fmaf(hv[i], wv[0], acc[0][i]);
                acc[1][i] = fmaf(hv[i], wv[1], acc[1][i]);
              }
            }
#pragma unroll
            for (int i = 0; i < SM_; ++i) {
              acc[0][i] += __shfl_xor_sync(0xffffffffu, acc[0][i], 8);
              acc[1][i] += __shfl_xor_sync(0xffffffffu, acc[1][i], 8);
              acc[0][i] += __shfl_xor_sync(0xffffffffu, acc[0][i], 16);
              acc[1][i] += __shfl_xor_sync(0xffffffffu, acc[1][i], 16);
            }
            if (cb == 0) {
              const int ud = S.reverse ? D - 1 - d : d;
              float uv[SM_], xv[SM_];
              ldv(uv, ubuf + ud * RP + r0);
#pragma unroll
              for (int i = 0; i < SM_; ++i) {
                const float mu = acc[0][i] + bo[0], lp = acc[1][i] + bo[1];
                const float mlp = fminf(-0.5f * lp, 10.f);             // jaxili/model.py:770
                ldinv[i] += mlp;                                        // 772 (last-pass values)
                xv[i] = mu + expf(mlp) * uv[i];                         // 771
              }
              stv(xbuf + d * RP + r0, xv);
            }
            __syncwarp();
          }
        }
        float* tmp = ubuf; ubuf = xbuf; xbuf = tmp;   // this layer's x is the next layer's u
      }
      __syncthreads();   // image buffer may be refilled
    }

    if (active) {   // theta = scale * x + shift (distrax.ScalarAffine.forward), coalesced rows
      float* dst = A.out + row0 * D;
      for (int idx = lane; idx < valid * D; idx += 32) {
        const int r = idx / D, j = idx - r * D;
        dst[idx] = c_scale[j] * ubuf[j * RP + r] + c_shift[j];
      }
      if (A.out_logdet && cb == 0) {
#pragma unroll
        for (int i = 0; i < SM_; ++i)
          if (r0 + i < valid) A.out_logdet[row0 + r0 + i] = ldinv[i];
      }
    }
    __syncwarp();
  }
}

// degree-sorted image: row/col positions come from spos (per layer, concatenated levels)
struct SamplePackArgs {
  int lvl_start[MAFB_MAX_LIN + 1];
  int spos_per_layer;
};
__global__ void pack_sample_kernel(const __grid_constant__ Plan P, const __grid_constant__ SamplePlan S,
                                   const __grid_constant__ SamplePackArgs SA,
                                   const float* __restrict__ params, const unsigned char* __restrict__ mask,
                                   const int* __restrict__ spos, float* __restrict__ simg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P.P) return;
  const int layer = idx / P.ppl;
  const int rem = idx - layer * P.ppl;
  int li = 0;
  for (int i = 0; i < P.n_lin; ++i)
    if (rem >= P.lin[i].pw_off) li = i;
  const LinDesc& ln = P.lin[li];
  const int* pos = spos + (size_t)layer * SA.spos_per_layer;
  float* base = simg + (size_t)layer * S.img_floats;
  const float val = mask[idx] ? params[idx] : 0.f;
  if (rem >= ln.pb_off) {
    const int n = rem - ln.pb_off;
    base[S.b_off[li] + pos[SA.lvl_start[li + 1] + n]] = val;
  } else {
    const int o = rem - ln.pw_off;
    const int k = o / ln.N, n = o - k * ln.N;
    base[S.w_off[li] + pos[SA.lvl_start[li] + k] * S.width[li + 1] + pos[SA.lvl_start[li + 1] + n]] = val;
  }
}

}  // namespace mafb
