// wide_kernels.cuh -- layer-by-layer ConditionalMAF path for models whose MADE does not fit shared
// memory (e.g. BASELINE config 5: n_in=32, n_cond=128, hidden [512,512], 10 layers).
//
// Activations live in HBM (row-major [rows][features]); every masked linear is a tiled GEMM with a
// fused epilogue (bias + activation, activation derivative, gradient masking).  The three product
// shapes of the MADE are one kernel template:
//   NN  C[m][n] = sum_k A[m][k]  W[k][n]      forward            h = act(a W' + b)
//   NT  C[m][k] = sum_n G[m][n]  W[k][n]      input gradient     g_in = (G W'^T) * act'
//   TN  C[k][n] = sum_m A[m][k]  G[m][n]      weight gradient    dW = M * (A^T G), split over rows
// W' is the mask-folded copy of the parameters (same flat layout), refreshed by the optimizer.
// This file is the FP32-FFMA core (exact parity with the oracle); the tcgen05 TF32 core replaces the
// mainloop for the large shapes (wide_tc.cuh).
#pragma once
#include <curand_kernel.h>
#include "common.cuh"

namespace mafb {

enum { EPI_BIAS_ACT = 0, EPI_BIAS = 1, EPI_MUL_DERIV = 2, EPI_ADD = 3, EPI_ATOMIC_MASK = 4 };

struct GemmArgs {
  const float* A; int lda;      // first operand
  const float* B; int ldb;      // second operand
  float* C; int ldc;            // output
  int M, N, K;                  // C is M x N, reduction length K
  const float* bias;            // EPI_BIAS*: [N]
  const float* aux; int ldaux;  // EPI_MUL_DERIV: h or act' [M x N]; EPI_ADD: addend [M x N]
  float* aux_out; int ldaux_out;  // EPI_BIAS_ACT with derivative output (non-ReLU): act'(a)
  const unsigned char* mask;    // EPI_ATOMIC_MASK: [M x N] (ld = N)
  int act;
  int k_chunk;                  // TN: reduction rows per CTA (split-K over blockIdx.z)
  float* colsum;                // optional [N]: column sums of the OUTPUT accumulate here (bias gradient
                                // of the previous linear, fused into the input-gradient product)
};

// 128 x BN x 8 tiles, 256 threads, 8 x (BN/16) outputs per thread, register-prefetched double buffer.
// TA: A tile is read with the reduction index as the fast (contiguous) global dimension (NN, NT);
//     otherwise the output-row index is contiguous (TN).
// TB: likewise for B (NT: reduction index contiguous; NN, TN: output-column index contiguous).
template <int BN, bool A_KFAST, bool B_KFAST, int EPI, bool RELU>
__global__ void __launch_bounds__(256) wide_gemm_kernel(const __grid_constant__ GemmArgs g) {
  constexpr int BM = 128, BK = 8, TN = BN / 16;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int k_begin = 0, k_end = g.K;
  if (EPI == EPI_ATOMIC_MASK) {
    k_begin = blockIdx.z * g.k_chunk;
    k_end = min(g.K, k_begin + g.k_chunk);
  }
  const int tm = (tid / 16) * 8, tn = (tid % 16) * TN;

  // global -> register staging: each thread moves one float4 of A (128x8 = 256 float4) and
  // BN*8/4/256 float4 of B per k-step
  float4 ra;
  float4 rb[(BN * BK / 4 + 255) / 256];
  auto load_tiles = [&](int k0) {
    if (A_KFAST) {   // A[m][k..k+3]: thread -> (m = tid/2, kq = tid%2)
      const int m = m0 + tid / 2, k = k0 + (tid % 2) * 4;
      ra = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < g.M) {
        const float* p = g.A + (size_t)m * g.lda + k;
        if (k + 3 < k_end) ra = *reinterpret_cast<const float4*>(p);
        else {
          if (k < k_end) ra.x = p[0];
          if (k + 1 < k_end) ra.y = p[1];
          if (k + 2 < k_end) ra.z = p[2];
        }
      }
    } else {         // A[k][m..m+3]: thread -> (k = tid/32, mq = tid%32)
      const int k = k0 + tid / 32, m = m0 + (tid % 32) * 4;
      ra = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < k_end) {
        const float* p = g.A + (size_t)k * g.lda + m;
        if (m + 3 < g.M) ra = *reinterpret_cast<const float4*>(p);
        else {
          if (m < g.M) ra.x = p[0];
          if (m + 1 < g.M) ra.y = p[1];
          if (m + 2 < g.M) ra.z = p[2];
        }
      }
    }
#pragma unroll
    for (int it = 0; it < (BN * BK / 4 + 255) / 256; ++it) {
      const int idx = tid + it * 256;
      rb[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < BN * BK / 4) {
        if (B_KFAST) {   // B[n][k..k+3]
          const int n = n0 + idx / 2, k = k0 + (idx % 2) * 4;
          if (n < g.N) {
            const float* p = g.B + (size_t)n * g.ldb + k;
            if (k + 3 < k_end) rb[it] = *reinterpret_cast<const float4*>(p);
            else {
              if (k < k_end) rb[it].x = p[0];
              if (k + 1 < k_end) rb[it].y = p[1];
              if (k + 2 < k_end) rb[it].z = p[2];
            }
          }
        } else {         // B[k][n..n+3]
          const int k = k0 + idx / (BN / 4), n = n0 + (idx % (BN / 4)) * 4;
          if (k < k_end) {
            const float* p = g.B + (size_t)k * g.ldb + n;
            if (n + 3 < g.N) rb[it] = *reinterpret_cast<const float4*>(p);
            else {
              if (n < g.N) rb[it].x = p[0];
              if (n + 1 < g.N) rb[it].y = p[1];
              if (n + 2 < g.N) rb[it].z = p[2];
            }
          }
        }
      }
    }
  };
  auto store_tiles = [&](int buf) {
    if (A_KFAST) {
      const int m = tid / 2, k = (tid % 2) * 4;
      As[buf][k][m] = ra.x; As[buf][k + 1][m] = ra.y; As[buf][k + 2][m] = ra.z; As[buf][k + 3][m] = ra.w;
    } else {
      const int k = tid / 32, m = (tid % 32) * 4;
      *reinterpret_cast<float4*>(&As[buf][k][m]) = ra;
    }
#pragma unroll
    for (int it = 0; it < (BN * BK / 4 + 255) / 256; ++it) {
      const int idx = tid + it * 256;
      if (idx < BN * BK / 4) {
        if (B_KFAST) {
          const int n = idx / 2, k = (idx % 2) * 4;
          Bs[buf][k][n] = rb[it].x; Bs[buf][k + 1][n] = rb[it].y; Bs[buf][k + 2][n] = rb[it].z; Bs[buf][k + 3][n] = rb[it].w;
        } else {
          const int k = idx / (BN / 4), n = (idx % (BN / 4)) * 4;
          *reinterpret_cast<float4*>(&Bs[buf][k][n]) = rb[it];
        }
      }
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (k_end - k_begin + BK - 1) / BK;
  if (nk > 0) {
    load_tiles(k_begin);
    store_tiles(0);
  }
  __syncthreads();
  for (int s = 0; s < nk; ++s) {
    const int buf = s & 1;
    if (s + 1 < nk) load_tiles(k_begin + (s + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[8], bv[TN];
      ldv(*reinterpret_cast<float(*)[4]>(&av[0]), &As[buf][kk][tm]);
      ldv(*reinterpret_cast<float(*)[4]>(&av[4]), &As[buf][kk][tm + 4]);
      if (TN == 8) {
        ldv(*reinterpret_cast<float(*)[4]>(&bv[0]), &Bs[buf][kk][tn]);
        ldv(*reinterpret_cast<float(*)[4]>(&bv[TN - 4]), &Bs[buf][kk][tn + TN - 4]);
      } else if (TN == 4) {
        ldv(*reinterpret_cast<float(*)[4]>(&bv[0]), &Bs[buf][kk][tn]);
      } else {
        ldv(*reinterpret_cast<float(*)[2]>(&bv[0]), &Bs[buf][kk][tn]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (s + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- fused epilogue
  float csum[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) csum[j] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + tm + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tn + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (EPI == EPI_BIAS_ACT) {
        float h, d;
        act_eval<RELU>(g.act, v + g.bias[n], h, d);
        g.C[(size_t)m * g.ldc + n] = h;
        if (!RELU && g.aux_out) g.aux_out[(size_t)m * g.ldaux_out + n] = d;
      } else if (EPI == EPI_BIAS) {
        g.C[(size_t)m * g.ldc + n] = v + g.bias[n];
      } else if (EPI == EPI_MUL_DERIV) {
        const float a = g.aux[(size_t)m * g.ldaux + n];
        const float o = RELU ? (a > 0.f ? v : 0.f) : v * a;
        g.C[(size_t)m * g.ldc + n] = o;
        if (g.colsum) csum[j] += o;
      } else if (EPI == EPI_ADD) {
        g.C[(size_t)m * g.ldc + n] = v + g.aux[(size_t)m * g.ldaux + n];
      } else {   // EPI_ATOMIC_MASK: masked weight gradient, rows split over blockIdx.z
        if (g.mask[(size_t)m * g.N + n]) atomicAdd(&g.C[(size_t)m * g.ldc + n], v);
      }
    }
  }
  if (EPI == EPI_MUL_DERIV && g.colsum) {
#pragma unroll
    for (int j = 0; j < TN; ++j)
      if (n0 + tn + j < g.N) atomicAdd(&g.colsum[n0 + tn + j], csum[j]);
  }
}

// ------------------------------------------------------------------ elementwise pieces
struct WideCommon {
  int D, C, K0, reverse;
  long long B;
  float inv_b, logdet_const;
};

// standardise x, y into the first layer's input [x' | y'] and copy y' into every layer's input
__global__ void wide_prep_kernel(WideCommon w, const float* __restrict__ x, const float* __restrict__ y,
                                 const float* __restrict__ stdc, float* __restrict__ a0, long long a0_stride,
                                 int L, float* __restrict__ ld) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w.B * w.K0) return;
  const long long r = idx / w.K0;
  const int j = (int)(idx - r * w.K0);
  if (j < w.D) {
    a0[r * w.K0 + j] = (x[r * w.D + j] - stdc[j]) * stdc[w.D + j];
    if (j == 0) ld[r] = 0.f;
  } else {
    const int c = j - w.D;
    const float v = __fdiv_rn(y[r * w.C + c] - stdc[2 * w.D + c], stdc[2 * w.D + w.C + c]);
    for (int l = 0; l < L; ++l) a0[l * a0_stride + r * w.K0 + j] = v;
  }
}

// affine autoregressive transform of one layer (jaxili/model.py:739-742): o = [mu | logp]
__global__ void wide_affine_kernel(WideCommon w, const float* __restrict__ o, const float* __restrict__ a0k,
                                   float* __restrict__ s_out, float* __restrict__ v_out,
                                   float* __restrict__ xnext, int xnext_ld, float* __restrict__ ld) {
  const long long r = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (r >= w.B) return;
  const int lane = threadIdx.x & 31;
  float lsum = 0.f;
  for (int j = lane; j < w.D; j += 32) {
    const float mu = o[r * 2 * w.D + j], lp = o[r * 2 * w.D + w.D + j];
    const float s = expf(0.5f * lp);
    const float v = (a0k[r * w.K0 + j] - mu) * s;
    if (s_out) { s_out[r * w.D + j] = s; v_out[r * w.D + j] = v; }
    const int jj = w.reverse ? w.D - 1 - j : j;
    xnext[r * xnext_ld + jj] = v;
    lsum += 0.5f * lp;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) ld[r] += lsum;
}

// log-probability per row, NLL accumulation, seed of the reverse pass (g_u = u / B)
__global__ void wide_final_kernel(WideCommon w, const float* __restrict__ u, const float* __restrict__ ld,
                                  float* __restrict__ out_logp, float* __restrict__ loss, float* __restrict__ gu) {
  const long long r = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  float contrib = 0.f;
  if (r < w.B) {
    float acc = 0.f;
    for (int j = lane; j < w.D; j += 32) {
      const float uv = u[r * w.D + j];
      acc += uv * uv;
      if (gu) gu[r * w.D + j] = uv * w.inv_b;
    }
    acc = warp_sum(acc);
    const float lp = -0.5f * acc - (float)w.D * 0.9189385332046727f + ld[r] + w.logdet_const;
    if (lane == 0) {
      if (out_logp) out_logp[r] = lp;
      contrib = -lp * w.inv_b;
    }
  }
  if (loss) {
    __shared__ float red[8];
    if (lane == 0) red[threadIdx.x / 32] = contrib;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < blockDim.x / 32; ++i) t += red[i];
      atomicAdd(loss, t);
    }
  }
}

// gradient w.r.t. the MADE outputs of one layer and the direct path to x
__global__ void wide_bwd_elem_kernel(WideCommon w, const float* __restrict__ gu, const float* __restrict__ s_in,
                                     const float* __restrict__ v_in, float* __restrict__ go,
                                     float* __restrict__ gxdir) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w.B * w.D) return;
  const long long r = idx / w.D;
  const int j = (int)(idx - r * w.D);
  const int jj = w.reverse ? w.D - 1 - j : j;
  const float gv = gu[r * w.D + jj];
  const float s = s_in[idx], v = v_in[idx];
  go[r * 2 * w.D + j] = -gv * s;
  go[r * 2 * w.D + w.D + j] = 0.5f * gv * v - 0.5f * w.inv_b;
  gxdir[idx] = gv * s;
}

// bias gradient: column sums of G [M x N] accumulated into db[N].  Block = 256 threads = 8 row lanes x 32 columns:
// a warp reads 32 consecutive columns of one row (coalesced), the 8 warps stride the block's rows, the partial
// sums meet in shared memory and leave as one reduction per column and block.
__global__ void __launch_bounds__(256) wide_colsum_kernel(const float* __restrict__ G, long long M, int N,
                                                          int rows_per_block, float* __restrict__ db) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (n < N) {
    long long r = r0 + w;
    for (; r + 24 < r1; r += 32) {     // four independent loads in flight per thread
      s0 += G[r * N + n];
      s1 += G[(r + 8) * N + n];
      s2 += G[(r + 16) * N + n];
      s3 += G[(r + 24) * N + n];
    }
    for (; r < r1; r += 8) s0 += G[r * N + n];
  }
  part[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][lane];
    atomicAdd(&db[n], t);
  }
}

// mask-folded copy of the parameters in the same flat layout (the wide path's "image")
__global__ void wide_pack_kernel(const float* __restrict__ params, const unsigned char* __restrict__ mask,
                                 float* __restrict__ wm, int P) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < P) wm[idx] = mask[idx] ? params[idx] : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// Inverse flow of the wide path (jaxili/model.py:745-773, 876-901, 923-947, 1116-1126): per MAF layer D
// sequential passes through the layer-by-layer MADE; pass d fixes x[:, d] = mu_d + exp(min(-logp_d / 2, 10)) u_d.

struct WideInvArgs {
  const float* u;                   // [N][D] base noise or nullptr -> Philox4x32-10, subsequence = global row
  unsigned long long seed, row_offset;
  const float* y_obs;               // [n_obs][C]
  long long n_obs, rows_per_obs;
  long long row0;                   // global index of this chunk's first row
};

// z <- base noise, a0 <- [0 | y'], log-det accumulator <- 0
__global__ void wide_inv_prep_kernel(WideCommon w, WideInvArgs a, const float* __restrict__ stdc,
                                     float* __restrict__ a0, float* __restrict__ z, float* __restrict__ ld) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= w.B) return;
  const int D = w.D, C = w.C;
  float* zr = z + r * D;
  if (a.u) {
    const float* src = a.u + r * D;
    for (int j = 0; j < D; ++j) zr[j] = src[j];
  } else {
    curandStatePhilox4_32_10_t st;
    curand_init(a.seed, a.row_offset + (unsigned long long)(a.row0 + r), 0ULL, &st);
    for (int j = 0; j < D; j += 4) {
      const float4 n4 = curand_normal4(&st);
      zr[j] = n4.x;
      if (j + 1 < D) zr[j + 1] = n4.y;
      if (j + 2 < D) zr[j + 2] = n4.z;
      if (j + 3 < D) zr[j + 3] = n4.w;
    }
  }
  long long obs = (a.row0 + r) / a.rows_per_obs;
  if (obs >= a.n_obs) obs = a.n_obs - 1;
  float* ar = a0 + r * w.K0;
  for (int j = 0; j < D; ++j) ar[j] = 0.f;
  for (int c = 0; c < C; ++c) ar[D + c] = __fdiv_rn(a.y_obs[obs * C + c] - stdc[2 * D + c], stdc[2 * D + C + c]);
  ld[r] = 0.f;
}

// pass d of one layer: x_d = mu_d + exp(m_d) * u_d, m = min(-logp / 2, 10); the last pass adds sum_j m_j to the
// log-determinant (every logp_j of that pass is final: it depends on x_<j only)
__global__ void wide_inv_update_kernel(WideCommon w, int d, const float* __restrict__ o, const float* __restrict__ z,
                                       float* __restrict__ a0, float* __restrict__ ld) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= w.B) return;
  const int D = w.D;
  const float* orow = o + r * 2 * D;
  const float m = fminf(-0.5f * orow[D + d], 10.f);
  const float zs = z[r * D + (w.reverse ? D - 1 - d : d)];      // u <- u[:, ::-1] before the passes (model.py:765)
  a0[r * w.K0 + d] = orow[d] + expf(m) * zs;
  if (d == D - 1) {
    float acc = 0.f;
    for (int j = 0; j < D; ++j) acc += fminf(-0.5f * orow[D + j], 10.f);
    ld[r] += acc;
  }
}

// layer done: its x becomes the next layer's u; the MADE input starts from zero again
__global__ void wide_inv_next_kernel(WideCommon w, float* __restrict__ a0, float* __restrict__ z) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w.B * w.D) return;
  const long long r = idx / w.D;
  const int j = (int)(idx - r * w.D);
  z[idx] = a0[r * w.K0 + j];
  a0[r * w.K0 + j] = 0.f;
}

// theta = scale * x + shift (distrax.ScalarAffine.forward, model.py:1122-1126)
__global__ void wide_inv_final_kernel(WideCommon w, const float* __restrict__ z, const float* __restrict__ ld,
                                      const float* __restrict__ stdc, float* __restrict__ out,
                                      float* __restrict__ out_logdet) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w.B * w.D) return;
  const long long r = idx / w.D;
  const int j = (int)(idx - r * w.D);
  out[idx] = stdc[2 * w.D + 2 * w.C + j] * z[idx] + stdc[j];
  if (out_logdet && j == 0) out_logdet[r] = ld[r];
}

// conditioner gradient of the wide path: sum over the layers of the y-columns of the first linear's input
// gradient, d/dy = (d/dy') / std
__global__ void wide_dy_accum_kernel(WideCommon w, const float* __restrict__ g0, float* __restrict__ acc, int first) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w.B * w.C) return;
  const long long r = idx / w.C;
  const int c = (int)(idx - r * w.C);
  const float v = g0[r * w.K0 + w.D + c];
  acc[idx] = first ? v : acc[idx] + v;
}
__global__ void wide_dy_final_kernel(WideCommon w, const float* __restrict__ acc, const float* __restrict__ stdc,
                                     float* __restrict__ out_dy) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w.B * w.C) return;
  const int c = (int)(idx % w.C);
  out_dy[idx] = __fdiv_rn(acc[idx], stdc[2 * w.D + w.C + c]);
}
// x-part of a [rows][K0] input gradient -> [rows][D] (+ the direct-path gradient)
__global__ void wide_dx_pick_kernel(WideCommon w, const float* __restrict__ g0, const float* __restrict__ gxdir,
                                    float* __restrict__ gu) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= w.B * w.D) return;
  const long long r = idx / w.D;
  const int j = (int)(idx - r * w.D);
  gu[idx] = g0[r * w.K0 + j] + gxdir[idx];
}

}  // namespace mafb
