// Micro-benchmark: the weight-gradient tile of the chain kernel's gradient warps (8x4 outputs per lane, 28 rows,
// 12 LDS.128 + 64 packed FFMA2 per 4 rows) with the kernel's lane -> tile mapping, NW warps on one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dw_tile dw_tile.cu && ./dw_tile
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void ldv(float (&d)[4], const float* p) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}
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

// MAP 0: the kernel's mapping (flat list over three linears: nKB x nNB = 3x13, 7x13, 7x6; kb = g / nNB, nb = g % nNB)
// MAP 1: all lanes the same tile (pure broadcast); MAP 2: kb = g % nKB, nb = g / nKB (lanes differ in kb first)
template <int MAP>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int Rt, int RP, int reps, float* pg = nullptr) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 2 * 64 * RP; i += blockDim.x) sm[i] = 1e-3f * (i & 15);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* Abase = sm;
  const float* Gbase = sm + 64 * RP;
  int g = warp * 32 + lane;
  int nKB, nNB, local;
  if (g < 39) { nKB = 3; nNB = 13; local = g; }
  else if (g < 130) { nKB = 7; nNB = 13; local = g - 39; }
  else { nKB = 7; nNB = 6; local = (g - 130) % 42; }
  int kb, nb;
  if (MAP == 0) { kb = local / nNB; nb = local - kb * nNB; }
  else if (MAP == 1) { kb = 0; nb = 0; }
  else { nb = local / nKB; kb = local - nb * nKB; }
  float o[8][4];
  float total = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < reps; ++it) {
    dw_tile(Abase + kb * RP, nKB * RP, Gbase + nb * RP, nNB * RP, Rt, o);
    if (pg) {   // the kernel's store phase: K = 50 (+ bias row), N = 50, this CTA's slice
      float* base = pg + (size_t)blockIdx.x * 23104 + (it & 3) * 4620;
      const int K = nKB == 3 ? 20 : 50, N = nNB == 6 ? 20 : 50;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kk = kb + i * nKB;
        if (kk <= K) {
          float* prow = base + kk * N;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = nb + j * nNB;
            if (n < N) prow[n] = o[i][j];
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) total += o[i][j];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = total;
  if (lane == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
  const int smem = 2 * 64 * 36 * 4 + 1024;
  const int reps = 100;
  for (int map = 0; map < 3; ++map)
    for (int nw : {1, 2, 4, 6, 8, 12}) {
      for (int rp : {36}) {
        long long h[16];
        for (int rep = 0; rep < 2; ++rep) {
          if (map == 0) k<0><<<148, nw * 32, smem>>>(out, cyc, 28, rp, reps);
          if (map == 1) k<1><<<148, nw * 32, smem>>>(out, cyc, 28, rp, reps);
          if (map == 2) k<2><<<148, nw * 32, smem>>>(out, cyc, 28, rp, reps);
          cudaDeviceSynchronize();
        }
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
        printf("map %d warps %2d RP %d: %.0f cycles per tile (7 iterations: %.0f per iteration)  err=%s\n", map, nw, rp,
               (double)mx / reps, (double)mx / reps / 7, cudaGetErrorString(cudaGetLastError()));
      }
    }
  float* pg; cudaMalloc(&pg, (size_t)148 * 23104 * 4);
  for (int nw : {1, 6}) {
    long long h[16];
    for (int rep = 0; rep < 2; ++rep) { k<0><<<148, nw * 32, smem>>>(out, cyc, 28, 36, reps, pg); cudaDeviceSynchronize(); }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("map 0 warps %2d with the store phase: %.0f cycles per tile  err=%s\n", nw, (double)mx / reps, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
