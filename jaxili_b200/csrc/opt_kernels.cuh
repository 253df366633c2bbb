// opt_kernels.cuh -- parameter packing, partial-gradient reduction, global-norm clip + Adam.
//
//   pack_kernel        mask-fold + pad + transpose the flat parameters into the per-layer
//                      shared-memory images the flow kernel bulk-copies
//                      (mask * w of jaxili/model.py:544, hoisted out of the step)
//   reduce_kernel      deterministic cross-CTA sum of the per-CTA partial gradients, masked
//                      (g_W = M * (h^T g)), plus loss and per-block squared-norm partials
//   sqnorm_kernel      squared-norm partials of an arbitrary (e.g. all-reduced) gradient
//   clip_adam_kernel   optax.chain(clip_by_global_norm, adam|sgd|adamw(warmup_cosine)) +
//                      apply_gradients (jaxili/train.py:246-300, 333), refreshing the images
#pragma once
#include "common.cuh"

namespace mafb {

struct OptArgs {
  float lr, warmup, decay_steps, init_factor, end_factor, b1, b2, eps, clip, weight_decay;
  int kind;
};

constexpr int RED_IDX = 128;    // parameters per reduction block (181 blocks for the narrow configs: every SM busy)
constexpr int RED_PARTS = 8;    // partial sums per parameter (part p takes the CTAs g = p, p + 8, ...)
constexpr int RED_THREADS = 32 * RED_PARTS;   // warp p = part p; lane q = parameters 4q .. 4q+3 of the block's 128
constexpr int RED_BATCH = 10;   // 128-bit loads in flight per thread

// fixed-order pairwise sum of the RED_PARTS partial sums of parameter i
__device__ __forceinline__ float red_parts_sum(const float (&red)[RED_PARTS][RED_IDX], int i) {
  static_assert(RED_PARTS == 8, "pairwise tree below is written for 8 parts");
  return ((red[0][i] + red[1][i]) + (red[2][i] + red[3][i])) + ((red[4][i] + red[5][i]) + (red[6][i] + red[7][i]));
}

// flat index -> (layer, linear, k, n); returns true for a bias entry
__device__ __forceinline__ bool decode_param(const Plan& P, int idx, int& layer, int& li, int& k, int& n) {
  layer = idx / P.ppl;
  const int rem = idx - layer * P.ppl;
  li = 0;
#pragma unroll 1
  for (int i = 0; i < P.n_lin; ++i)
    if (rem >= P.lin[i].pw_off) li = i;
  const LinDesc& ln = P.lin[li];
  if (rem >= ln.pb_off) {
    k = 0;
    n = rem - ln.pb_off;
    return true;
  }
  const int o = rem - ln.pw_off;
  k = o / ln.N;
  n = o - k * ln.N;
  return false;
}

// where parameter idx lives in the weight images (float offsets from the image base): o1 = forward image,
// o2 = transposed (reverse-pass) image, -1 for a bias
__device__ __forceinline__ void image_offsets(const Plan& P, int idx, long long& o1, long long& o2) {
  int layer, li, k, n;
  const bool is_bias = decode_param(P, idx, layer, li, k, n);
  const LinDesc& ln = P.lin[li];
  const long long base = (long long)layer * P.img_floats;
  // forward image of the LAST linear: output columns interleaved (mu_j, logp_j) -> (2j, 2j+1), so that one
  // thread's 4-column tile holds both halves of a feature and the affine transform runs in the epilogue.
  // The transposed (reverse-pass) image keeps the parameter order.
  const int nf = (li == P.n_lin - 1) ? (n < P.D ? 2 * n : 2 * (n - P.D) + 1) : n;
  if (is_bias) {
    o1 = base + ln.b_off + nf;
    o2 = -1;
  } else {
    o1 = base + ln.w_off + k * ln.NP + nf;
    o2 = base + P.fwd_floats + ln.wt_off + n * ln.KP + k;
  }
}

__device__ __forceinline__ void write_image(const Plan& P, float* wimg, int idx, float masked_value) {
  long long o1, o2;
  image_offsets(P, idx, o1, o2);
  wimg[o1] = masked_value;
  if (o2 >= 0) wimg[o2] = masked_value;
}

__global__ void pack_kernel(const __grid_constant__ Plan P, const float* __restrict__ params,
                            const unsigned char* __restrict__ mask, float* __restrict__ wimg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P.P) return;
  write_image(P, wimg, idx, mask[idx] ? params[idx] : 0.f);
}

// Part sums of the block's 128 parameters over the per-CTA partial gradients [G][stride]: warp p adds the rows
// g = p, p + 8, ... in that order, four neighbouring parameters per lane with one 128-bit load per row (a warp reads
// 512 contiguous bytes), RED_BATCH rows requested before the first add.  A step reads G * P * 4 bytes here (13.7 MB
// at c2) out of L2: the loads in flight, not the adds, set the time.
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partial, int stride, int G, int P,
                                                float (&red)[RED_PARTS][RED_IDX]) {
  const int q = threadIdx.x & 31, p = threadIdx.x >> 5;
  const int i4 = blockIdx.x * RED_IDX + 4 * q;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i4 < P) {                                       // stride % 4 == 0 and stride >= P: the 128-bit read stays in the row
    const float4* src = reinterpret_cast<const float4*>(partial + i4);
    const size_t st4 = (size_t)(stride >> 2);
    for (int g0 = p; g0 < G; g0 += RED_PARTS * RED_BATCH) {
      float4 v[RED_BATCH];
#pragma unroll
      for (int u = 0; u < RED_BATCH; ++u) {
        const int g = g0 + u * RED_PARTS;
        if (g < G) v[u] = __ldcg(src + (size_t)g * st4);
      }
#pragma unroll
      for (int u = 0; u < RED_BATCH; ++u)
        if (g0 + u * RED_PARTS < G) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
  }
  *reinterpret_cast<float4*>(&red[p][4 * q]) = s;
}

// block: RED_THREADS threads.  normpart[blockIdx] = sum of grad^2 over the block.
__global__ void __launch_bounds__(RED_THREADS)
reduce_kernel(const float* __restrict__ partial, int stride, int G, int P, const unsigned char* __restrict__ mask,
              float* __restrict__ grad, const float* __restrict__ partial_loss, float* __restrict__ loss,
              float* __restrict__ normpart, const int* __restrict__ dp_step, float* dp_slot0,
              float* dp_slot1) {
  if (dp_step) {
    // data parallel over peer memory: the reduced gradient goes to this rank's symmetric slot of the
    // coming epoch (= optax count + 1, a device counter, so the step stays CUDA-graph capturable)
    grad = ((dp_step[0] + 1) & 1) ? dp_slot1 : dp_slot0;
    loss = grad + ((P + 3) & ~3);
  }
  __shared__ __align__(16) float red[RED_PARTS][RED_IDX];
  __shared__ float wsum[RED_IDX / 32];
  const int i = threadIdx.x % RED_IDX, p = threadIdx.x / RED_IDX;   // after the part sums: one parameter per thread of group 0
  const int idx = blockIdx.x * RED_IDX + i;
  reduce_partials(partial, stride, G, P, red);
  __syncthreads();
  if (p == 0) {
    float tot = 0.f;
    if (idx < P) {
      tot = red_parts_sum(red, i);
      tot = mask[idx] ? tot : 0.f;
      grad[idx] = tot;
    }
    const float sq = warp_sum(tot * tot);
    if ((i & 31) == 0) wsum[i >> 5] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < RED_IDX / 32; ++w) t += wsum[w];
    normpart[blockIdx.x] = t;
  }
  if (blockIdx.x == 0 && p == 1) {   // second part-group of block 0 reduces the loss
    float l = 0.f;
    for (int g = i; g < G; g += RED_IDX) l += partial_loss[g];
    l = warp_sum(l);
    __shared__ float lsum[RED_IDX / 32];
    if ((i & 31) == 0) lsum[i >> 5] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(RED_IDX));
    if (i == 0) {
      float t = 0.f;
      for (int w = 0; w < RED_IDX / 32; ++w) t += lsum[w];
      loss[0] = t;
    }
  }
}

__global__ void __launch_bounds__(RED_IDX)
sqnorm_kernel(const float* __restrict__ grad, int P, float* __restrict__ normpart) {
  __shared__ float wsum[RED_IDX / 32];
  const int idx = blockIdx.x * RED_IDX + threadIdx.x;
  const float g = idx < P ? grad[idx] : 0.f;
  const float sq = warp_sum(g * g);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < RED_IDX / 32; ++w) t += wsum[w];
    normpart[blockIdx.x] = t;
  }
}

// sum of the squared-norm partials in a fixed order (one block): large models have tens of thousands of partials,
// which every block of clip_adam_kernel would otherwise re-read
__global__ void __launch_bounds__(1024) sum_partials_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
  __shared__ float wsum[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) s += part[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += wsum[w];
    out[0] = t;
  }
}

// optax.warmup_cosine_decay_schedule evaluated in FP32 at integer count t
__device__ __forceinline__ float lr_schedule(const OptArgs& o, int t) {
  const float tf = (float)t;
  const float peak = o.lr, init = o.init_factor * o.lr;
  if (tf < o.warmup) {
    const float frac = 1.f - tf / o.warmup;
    return (init - peak) * frac + peak;
  }
  const float dec = o.decay_steps - o.warmup;
  const float tau = fminf(tf - o.warmup, dec);
  const float cosv = 0.5f * (1.f + cosf(3.14159265358979323846f * tau / dec));
  const float alpha = o.end_factor;
  return peak * ((1.f - alpha) * cosv + alpha);
}

// grid: ceil(P / RED_IDX) blocks of RED_IDX threads; n_norm = number of norm partials.
__global__ void __launch_bounds__(RED_IDX)
clip_adam_kernel(const __grid_constant__ Plan P, const __grid_constant__ OptArgs O,
                 float* __restrict__ params, const float* __restrict__ grad, float* __restrict__ m,
                 float* __restrict__ v, int* __restrict__ step, const float* __restrict__ normpart,
                 int n_norm, const unsigned char* __restrict__ mask, float* __restrict__ wimg, int flat_image,
                 unsigned int* __restrict__ done_counter, float* __restrict__ gnorm_out) {
  __shared__ float wsum[RED_IDX / 32];
  __shared__ float s_norm;
  const int t = step[0];
  // every block sums the norm partials in the same fixed order -> identical scale everywhere
  float s = 0.f;
  for (int i = threadIdx.x; i < n_norm; i += RED_IDX) s += normpart[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < RED_IDX / 32; ++w) tot += wsum[w];
    s_norm = sqrtf(tot);
  }
  __syncthreads();
  const float gnorm = s_norm;
  const int idx = blockIdx.x * RED_IDX + threadIdx.x;
  if (idx < P.P) {
    float g = grad[idx];
    // optax.clip_by_global_norm: g if norm < max_norm else (g / norm) * max_norm
    if (!(gnorm < O.clip)) g = (g / gnorm) * O.clip;
    float p = params[idx];
    const float lr_t = lr_schedule(O, t);
    float upd;
    if (O.kind == 1) {  // sgd (+ add_decayed_weights before it)
      upd = g + O.weight_decay * p;
    } else {
      const float mm = O.b1 * m[idx] + (1.f - O.b1) * g;
      const float vv = O.b2 * v[idx] + (1.f - O.b2) * g * g;
      m[idx] = mm;
      v[idx] = vv;
      const float tn = (float)(t + 1);
      const float mhat = mm / (1.f - powf(O.b1, tn));
      const float vhat = vv / (1.f - powf(O.b2, tn));
      upd = mhat / (sqrtf(vhat) + O.eps);
      if (O.kind == 2) upd += O.weight_decay * p;  // adamw
    }
    p = p - lr_t * upd;
    params[idx] = p;
    if (flat_image) wimg[idx] = mask[idx] ? p : 0.f;      // wide path: masked copy, same layout
    else write_image(P, wimg, idx, mask[idx] ? p : 0.f);
  }
  // the last block to finish bumps the optax count (all blocks have read step[0] by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) {
      step[0] = t + 1;
      *done_counter = 0u;
      if (gnorm_out) gnorm_out[0] = gnorm;
    }
  }
}

// Validation / test loop on the device (jaxili/train.py:453-487: size-weighted mean of the batch losses =
// -sum(log_prob) / rows): acc[0] += -sum_b lp[b], acc[1] += B, one block, fixed summation order.
__global__ void __launch_bounds__(1024) eval_sum_kernel(const float* __restrict__ lp, long long B, float* __restrict__ acc) {
  __shared__ float wsum[32];
  float s = 0.f;
  for (long long i = threadIdx.x; i < B; i += 1024) s -= lp[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += wsum[w];
    acc[0] += t;
    acc[1] += (float)B;
  }
}

// FP32 FFMA throughput probe with the operand pattern of the register-tiled GEMM inner loops:
// acc[j][i] = fma(a[i], w[j], acc[j][i]) on 16 independent accumulators, all-register operands.
__global__ void ffma_peak_kernel(float* out, int iters, float seed) {
  float a[4], w[4], acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = seed * (float)(threadIdx.x + i + 1);
    w[i] = seed * (float)(threadIdx.x * 3 + i + 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][i] = (float)(i + j);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(a[i], w[j], acc[j][i]);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += acc[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same outer product with packed FFMA2 (fma.rn.f32x2, what the flow kernels issue): 8 per 4x4 step
__global__ void ffma2_peak_kernel(float* out, int iters, float seed) {
  float a[4], w[4], acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = seed * (float)(threadIdx.x + i + 1);
    w[i] = seed * (float)(threadIdx.x * 3 + i + 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][i] = (float)(i + j);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; i += 2) ffma2(acc[j][i], acc[j][i + 1], a[i], a[i + 1], w[j], w[j]);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += acc[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace mafb
