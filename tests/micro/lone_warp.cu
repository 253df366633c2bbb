// Micro-benchmark: how fast can ONE warp per SM sub-partition run the chain kernel's inner loop?
// (16 packed FFMA2 + 3 LDS.128 per step, operands through a 4-stage register ring.)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lone_warp lone_warp.cu && ./lone_warp
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void tile(float (&acc)[4][4], const float4& a, const float4& w, int mode) {
  const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
  if (mode == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(av[i], wv[j], acc[j][i]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; i += 2) ffma2(acc[j][i], acc[j][i + 1], av[i], av[i + 1], wv[j], wv[j]);
  }
}

// MODE 0: FFMA2, operands fixed in registers; 1: FFMA2 + ring-prefetched LDS.128 x3 per step; 2: as 1 with plain FFMA
// 3: as 1 plus a 16-shuffle reduce every 26 steps
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int steps) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3f * (i & 15);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[2][4][4];
#pragma unroll
  for (int g = 0; g < 2; ++g)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[g][j][i] = 0.f;
  const float* pa = sm + warp * 8 + (lane >> 4) * 36;
  const float* pw = sm + 4096 + (lane & 15) * 4 + (lane >> 4) * 52;
  const int da = 72, dw = 104;
  float4 a0[4], a1[4], w[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    a0[u] = *(const float4*)(pa + u * da);
    a1[u] = *(const float4*)(pa + u * da + 4);
    w[u] = *(const float4*)(pw + u * dw);
  }
  const long long t0 = clock64();
#pragma unroll 1
  for (int s = 0; s < steps; s += 4) {
    const int o = (s & 31);   // stay inside the buffer
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      tile(acc[0], a0[u], w[u], MODE);
      tile(acc[1], a1[u], w[u], MODE);
      if (MODE >= 1) {
        a0[u] = *(const float4*)(pa + (o + u) * da % 2048);
        a1[u] = *(const float4*)(pa + (o + u) * da % 2048 + 4);
        w[u] = *(const float4*)(pw + (o + u) * dw % 2048);
      }
    }
    if (MODE == 3 && (s % 24) == 20) {
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool hi = lane & 16;
            const float send = hi ? acc[g][c][i] : acc[g][2 + c][i];
            const float mine = hi ? acc[g][2 + c][i] : acc[g][c][i];
            acc[g][c][i] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
          }
    }
  }
  const long long t1 = clock64();
  float sum = 0.f;
#pragma unroll
  for (int g = 0; g < 2; ++g)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) sum += acc[g][j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, int grid) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  const int steps = 4096;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<MODE><<<grid, warps * 32, 65536>>>(out, cyc, steps);
  k<MODE><<<grid, warps * 32, 65536>>>(out, cyc, steps);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/CTA %2d: %.1f cycles per step per warp (16 FFMA2-equivalents; FMA-pipe bound = %.0f)\n", name, warps,
         (double)c / steps, 32.0 * ((warps + 3) / 4));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) run<0>("FFMA2 regs only", w, 148);
  for (int w : {4, 8, 16}) run<1>("FFMA2 + 3 LDS.128 ring", w, 148);
  for (int w : {4, 8, 16}) run<2>("FFMA + 3 LDS.128 ring", w, 148);
  for (int w : {4, 8}) run<3>("FFMA2 + LDS + shuffle reduce", w, 148);
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
