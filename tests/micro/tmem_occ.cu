// Does a kernel that allocates tensor memory (tcgen05.alloc) get more than one CTA per SM?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(192) k_plain(float* out) { out[blockIdx.x * blockDim.x + threadIdx.x] = 1.f; }
__global__ void __launch_bounds__(192) k_tmem(float* out, unsigned* slot_out) {
  __shared__ unsigned slot;
  if (threadIdx.x < 32) {
    unsigned a = (unsigned)__cvta_generic_to_shared(&slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "n"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  unsigned base = slot;
  // hold the SM for a while so co-residency shows up in the %smid histogram
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {}
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) slot_out[blockIdx.x] = smid;
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)base;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(32));
}
int main() {
  int o1 = 0, o2 = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_plain, 192, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_tmem, 192, 0);
  printf("occupancy API: plain %d, tcgen05.alloc kernel %d\n", o1, o2);
  float* out; unsigned* sm; cudaMalloc(&out, 296 * 192 * 4); cudaMalloc(&sm, 296 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid : {148, 296}) {
    cudaEventRecord(e0); k_tmem<<<grid, 192>>>(out, sm); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("grid %d: %.3f ms (%s)\n", grid, ms, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
