// dp_kernel.cuh -- one-shot all-reduce of the flat gradient over NVLink peer memory, fused with the
// squared-norm partials the clip needs.
//
// Data-parallel training exchanges P floats (+ the loss) per step: 92 KB for the narrow configs.  At
// that size a ring/tree collective is pure latency (NCCL: ~30 us per step against a 70 us step on
// 8xB200).  Here every rank keeps its locally reduced gradient in a symmetric buffer that all peers
// map (CUDA IPC); one kernel per rank (1) raises its epoch flag in every peer's flag array
// (release, system scope), (2) waits until all peers have raised theirs (acquire), (3) sums the
// world's buffers with plain peer loads in a FIXED rank order -- so all replicas get bit-identical
// gradients -- writes the global gradient + loss locally and emits the norm partials.
// Buffers alternate with the epoch parity: a rank can only reach epoch e+2 (and overwrite slot e&1)
// after every peer has signalled e+1, i.e. has finished reading slot e&1.
//
// One such kernel runs per GPU (one process per GPU); the spin has a time-out that sets an error
// word instead of hanging the device.
#pragma once
#include "common.cuh"
#include "opt_kernels.cuh"

namespace mafb {

constexpr int DP_MAX_WORLD = 8;

struct DpArgs {
  int rank, world, P;
  const int* step;                        // device optax count; epoch = step + 1, slot parity = epoch & 1
  const float* sym[DP_MAX_WORLD];         // peers' (and own) symmetric buffers: 2 slots of slot_floats, loss at [r4(P)]
  long long slot_floats;
  unsigned int* flags_of[DP_MAX_WORLD];   // flag array [DP_MAX_WORLD] living on rank r
  float* grad_out;                        // local [P]
  float* loss_out;                        // local [1]
  float* normpart;                        // local [ceil(P / RED_IDX)]
  unsigned int* error;                    // local: set to 1 on time-out
  long long timeout_cycles;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// grid: ceil(P / RED_IDX) blocks of RED_IDX threads (same partition as sqnorm_kernel / clip_adam_kernel)
__global__ void __launch_bounds__(RED_IDX) dp_allreduce_kernel(const __grid_constant__ DpArgs a) {
  __shared__ float wsum[RED_IDX / 32];
  __shared__ int s_ok;
  const int tid = threadIdx.x;
  const unsigned int epoch = (unsigned int)a.step[0] + 1u;
  const long long soff = (long long)(epoch & 1u) * a.slot_floats;
  if (blockIdx.x == 0 && tid < a.world) {
    __threadfence_system();                                  // this rank's slot (previous kernel) is visible
    st_release_sys(a.flags_of[tid] + a.rank, epoch);          // tell rank `tid`
  }
  if (tid == 0) s_ok = 1;
  __syncthreads();
  if (tid < a.world) {
    const unsigned int* f = a.flags_of[a.rank] + tid;         // my flag array, entry of rank `tid`
    const long long t0 = clock64();
    // epochs only grow; signed difference tolerates wrap-around
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {
      if (clock64() - t0 > a.timeout_cycles) {
        s_ok = 0;
        *a.error = 1u;
        break;
      }
    }
  }
  __syncthreads();
  const int idx = blockIdx.x * RED_IDX + tid;
  float g = 0.f;
  if (s_ok && idx < a.P) {
#pragma unroll
    for (int r = 0; r < DP_MAX_WORLD; ++r)
      if (r < a.world) g += a.sym[r][soff + idx];             // fixed order: replicas stay bit-identical
    a.grad_out[idx] = g;
  }
  const float sq = warp_sum(g * g);
  if ((tid & 31) == 0) wsum[tid >> 5] = sq;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < RED_IDX / 32; ++w) t += wsum[w];
    a.normpart[blockIdx.x] = t;
    if (blockIdx.x == 0) {
      const int lo = (a.P + 3) & ~3;
      float l = 0.f;
      for (int r = 0; r < a.world; ++r) l += a.sym[r][soff + lo];
      a.loss_out[0] = s_ok ? l : __int_as_float(0x7fc00000);
    }
  }
}

}  // namespace mafb
