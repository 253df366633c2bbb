// FFMA vs packed FFMA2 (fma.rn.f32x2) issue-rate probe for sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_ffma(float* out, int iters, float seed) {
  float a[4], w[4], acc[4][4];
  for (int i = 0; i < 4; ++i) { a[i] = seed * (threadIdx.x + i + 1); w[i] = seed * (threadIdx.x * 3 + i + 1);
    for (int j = 0; j < 4; ++j) acc[j][i] = i + j; }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(a[i], w[j], acc[j][i]);
  float s = 0; for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) s += acc[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack(float x, float y) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}

// same 4x4 outer product per step: 8 FFMA2 (acc pairs along i), w[j] duplicated into both halves ONCE (outside)
__global__ void k_ffma2(float* out, int iters, float seed) {
  unsigned long long a2[2], w2[4], acc[4][2];
  for (int i = 0; i < 2; ++i) a2[i] = pack(seed * (threadIdx.x + 2 * i + 1), seed * (threadIdx.x + 2 * i + 2));
  for (int j = 0; j < 4; ++j) { float w = seed * (threadIdx.x * 3 + j + 1); w2[j] = pack(w, w);
    for (int i = 0; i < 2; ++i) acc[j][i] = pack(i + j, i + j + 1); }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) fma2(acc[j][i], a2[i], w2[j]);
  unsigned long long s = 0; for (int j = 0; j < 4; ++j) for (int i = 0; i < 2; ++i) s ^= acc[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}

// with the duplication inside the loop: 4 packs + 8 FFMA2 per 4x4 step (what a real smem-fed loop would pay)
__global__ void k_ffma2_dup(float* out, int iters, float seed) {
  unsigned long long a2[2], acc[4][2];
  float w[4];
  for (int i = 0; i < 2; ++i) a2[i] = pack(seed * (threadIdx.x + 2 * i + 1), seed * (threadIdx.x + 2 * i + 2));
  for (int j = 0; j < 4; ++j) { w[j] = seed * (threadIdx.x * 3 + j + 1);
    for (int i = 0; i < 2; ++i) acc[j][i] = pack(i + j, i + j + 1); }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float wj = w[j] + (float)u;                    // defeat hoisting of the pack
        unsigned long long w2 = pack(wj, wj);
#pragma unroll
        for (int i = 0; i < 2; ++i) fma2(acc[j][i], a2[i], w2);
      }
  unsigned long long s = 0; for (int j = 0; j < 4; ++j) for (int i = 0; i < 2; ++i) s ^= acc[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}

template <typename F> float run(F f, const char* name, int threads) {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  int iters = 20000, grid = 148 * (2048 / threads);
  f<<<grid, threads>>>(out, 10, 1e-9f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); f<<<grid, threads>>>(out, iters, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flop = 2.0 * 128 * iters * (double)grid * threads;
  printf("%-12s threads/CTA %4d  %.1f TFLOP/s (%s)\n", name, threads, flop / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); return ms;
}
int main() {
  for (int t : {256, 512, 1024}) { run(k_ffma, "FFMA", t); run(k_ffma2, "FFMA2", t); run(k_ffma2_dup, "FFMA2+dup", t); }
  return 0;
}
