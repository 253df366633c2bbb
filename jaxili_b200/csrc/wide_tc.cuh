// wide_tc.cuh -- tcgen05 TF32 GEMM core of the wide ConditionalMAF path (sm_100a).
//
//   C[M x N] (+)= op(A) . op(B),  128 x BN output tile per CTA, FP32 accumulation in TMEM.
//
// Pipeline (one tile per CTA, 2 CTAs per SM so one CTA's epilogue overlaps the other's mainloop):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B) into a STAGES-deep smem ring
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma.kind::tf32
//               (M=128, N=BN, K=8), 4 per 32-wide k-block; tcgen05.commit frees the smem slot
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> fused epilogue -> global
//
// Operand layouts.  "K-major": the reduction index is contiguous in global memory; a k-block is one
// TMA box [rows x 32 floats] = rows x 128 B, the canonical SW128 K-major UMMA layout (8-row groups
// 1024 B apart).  "MN-major": the output index is contiguous; a k-block is a column of 32x32 boxes
// (32 k-rows x 128 B each, 4 KB apart, TMA swizzle 128B_ATOM_32B), the canonical
// SWIZZLE_128B_BASE32B MN-major layout.  Both are legal for
// kind::tf32, which is what lets all three products of the MADE run without transposed copies:
//   forward      h = act(a W' + b):        A = a   [M][K]  K-major,   B = W' [K][N] MN-major
//   input grad   g = (G W'^T) * act':      A = G   [M][N'] K-major,   B = W' [K'][N'] K-major
//   weight grad  dW = M * (a^T G):         A = a   [m][K'] MN-major,  B = G  [m][N]  MN-major (split over m)
// Encodings follow cute/arch/mma_sm100_desc.hpp (InstrDescriptor, SmemDescriptor).
#pragma once
#include <cuda.h>
#include "wide_kernels.cuh"

namespace mafb {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;            // floats per k-block = one 128-byte swizzle row

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (cute SmemDescriptor: version 1 for sm_100).
// layout_type 2 = SWIZZLE_128B (K-major operands); 1 = SWIZZLE_128B_BASE32B, the only layout the
// tensor core accepts for MN-major 32-bit operands (32-byte swizzle atoms, 4-row K groups).
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                 uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;      // version_ = 1
  d |= (uint64_t)layout_type << 61;
  return d;
}

// instruction descriptor for kind::tf32, FP32 accumulate (cute InstrDescriptor)
__host__ __device__ constexpr uint32_t tc_instr_desc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4)                       // c_format = F32
         | (2u << 7) | (2u << 10)        // a_format = b_format = TF32
         | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BM, int BN>
struct TcCfg {
  static constexpr int MH = BM / 128;                   // 128-row halves, one MMA + accumulator each
  // 128x256: two stages (97 KB) so that TWO CTAs share an SM (256 TMEM columns each): one CTA's epilogue
  // (TMEM -> registers -> shared -> global) overlaps the other's TMA/MMA main loop
  static constexpr int STAGES = (BM == 128 && BN == 256) ? 2 : ((BM * BN >= 256 * 256) ? 3 : (BN >= 128 ? 3 : 4));
  static constexpr int A_BYTES = BM * TC_BK * 4;        // 16 KB per half
  static constexpr int B_BYTES = BN * TC_BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int THREADS = 64 + 128 * MH;         // TMA warp, MMA warp, 4 epilogue warps per half
  static constexpr int TMEM_COLS = MH * BN;             // power of two in {32, ..., 512}
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
};

// Tile BM x BN per CTA, BM in {128, 256}.  BM = 256 runs two M=128 MMAs per k-step against the same
// B tile (32 MAC per operand byte instead of 16: with 4-byte TF32 operands the 128x128 tile is
// L2-bandwidth-bound), its two accumulators fill all 512 TMEM columns.
template <int BM, int BN, bool A_MN, bool B_MN, int EPI, bool RELU>
__global__ void __launch_bounds__(TcCfg<BM, BN>::THREADS)
wide_tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ GemmArgs g) {
  using Cfg = TcCfg<BM, BN>;
  constexpr int STAGES = Cfg::STAGES, MH = Cfg::MH;
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* accum_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int k_begin = 0, k_end = g.K;
  if (EPI == EPI_ATOMIC_MASK) {
    k_begin = blockIdx.z * g.k_chunk;
    k_end = min(g.K, k_begin + g.k_chunk);
  }
  const int num_kb = (k_end - k_begin + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {   // TMEM: MH accumulators of BN FP32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        if (kb >= STAGES) mbar_wait(&empty[s], (uint32_t)(((kb / STAGES) - 1) & 1));
        unsigned char* a_dst = tiles + s * Cfg::STAGE_BYTES;
        unsigned char* b_dst = a_dst + Cfg::A_BYTES;
        const int k0 = k_begin + kb * TC_BK;
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        if (A_MN) {
#pragma unroll
          for (int a = 0; a < BM / 32; ++a) tma_load_2d(a_dst + a * 4096, &tmA, m0 + 32 * a, k0, &full[s]);
        } else {
#pragma unroll
          for (int hh = 0; hh < MH; ++hh) tma_load_2d(a_dst + hh * 16384, &tmA, k0, m0 + 128 * hh, &full[s]);
        }
        if (B_MN) {
#pragma unroll
          for (int a = 0; a < BN / 32; ++a) tma_load_2d(b_dst + a * 4096, &tmB, n0 + 32 * a, k0, &full[s]);
        } else {
          tma_load_2d(b_dst, &tmB, k0, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc_instr_desc(128, BN, A_MN, B_MN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full[s], (uint32_t)((kb / STAGES) & 1));
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + s * Cfg::STAGE_BYTES);
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          // K-major: step 8 floats (32 B) inside the swizzled 128-B row; 8-row groups 1024 B apart.
          // MN-major: step 8 k-rows (1024 B) = two 4-row groups 512 B apart; 32-wide MN atoms 4096 B apart.
          const uint64_t bd = B_MN ? tc_smem_desc(b_addr + k * 1024, 4096, 512, 1)
                                   : tc_smem_desc(b_addr + k * 32, 16, 1024, 2);
#pragma unroll
          for (int hh = 0; hh < MH; ++hh) {
            const uint32_t ah = a_addr + hh * 16384;   // 128 rows of A: 16 KB in either layout
            const uint64_t ad = A_MN ? tc_smem_desc(ah + k * 1024, 4096, 512, 1)
                                     : tc_smem_desc(ah + k * 32, 16, 1024, 2);
            tc_mma_tf32(tmem_base + hh * BN, ad, bd, idesc, (kb | k) ? 1u : 0u);
          }
        }
        tc_commit(&empty[s]);          // smem slot reusable once these MMAs have read it
      }
      tc_commit(accum_full);           // accumulators complete
    }
  } else {
    // ===== epilogue warps: TMEM lane group = warp % 4, accumulator half = (warp - 2) / 4.
    // tcgen05.ld hands every lane one accumulator ROW; each 32x32 chunk is transposed through shared
    // memory (the pipeline buffers are drained once the accumulator barrier fires) so that global
    // loads / stores / reductions are coalesced 128-byte row segments.
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int lg = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int m_base = m0 + hh * 128 + lg * 32;
    float* stg = reinterpret_cast<float*>(tiles) + (warp - 2) * (32 * 33);
    float* csm = reinterpret_cast<float*>(tiles) + 8 * (32 * 33);      // [4 * MH warps][BN] column sums
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float csum = 0.f;
      float v[32];
      if (num_kb > 0) tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + hh * BN + c, v);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
      __syncwarp();
      const int n = n0 + c + lane;
      if (n < g.N) {
        float bn = 0.f;
        if (EPI == EPI_BIAS_ACT || EPI == EPI_BIAS) bn = g.bias[n];
        const int rows = min(32, g.M - m_base);
        // all 32 row-segment loads of the chunk are issued before any is used (memory-level parallelism)
        float a[32];
        if (EPI == EPI_MUL_DERIV || EPI == EPI_ADD) {
          const float* ap = g.aux + (size_t)m_base * g.ldaux + n;
#pragma unroll
          for (int r = 0; r < 32; ++r) a[r] = r < rows ? ap[(size_t)r * g.ldaux] : 0.f;
        }
        if (EPI == EPI_ATOMIC_MASK) {
          const unsigned char* mp = g.mask + (size_t)m_base * g.N + n;
#pragma unroll
          for (int r = 0; r < 32; ++r) a[r] = (r < rows && mp[(size_t)r * g.N]) ? 1.f : 0.f;
        }
        float* cp = g.C + (size_t)m_base * g.ldc + n;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          float val = stg[r * 33 + lane];
          if (EPI == EPI_ATOMIC_MASK) {
            if (a[r] != 0.f)
              asm volatile("red.global.add.f32 [%0], %1;" ::"l"(cp + (size_t)r * g.ldc), "f"(val) : "memory");
          } else if (r < rows) {
            if (EPI == EPI_BIAS_ACT) {
              float hv, dd;
              act_eval<RELU>(g.act, val + bn, hv, dd);
              val = hv;
              if (!RELU && g.aux_out) g.aux_out[(size_t)(m_base + r) * g.ldaux_out + n] = dd;
            } else if (EPI == EPI_BIAS) {
              val = val + bn;
            } else if (EPI == EPI_MUL_DERIV) {
              val = RELU ? (a[r] > 0.f ? val : 0.f) : val * a[r];
              csum += val;
            } else if (EPI == EPI_ADD) {
              val = val + a[r];
            }
            cp[(size_t)r * g.ldc] = val;
          }
        }
      }
      if (EPI == EPI_MUL_DERIV && g.colsum) csm[(warp - 2) * BN + c + lane] = csum;
      __syncwarp();
    }
    if (EPI == EPI_MUL_DERIV && g.colsum) {
      // bias gradient of the previous linear = column sums of this output: combine the CTA's
      // epilogue warps in shared memory, then one coalesced reduction per column
      asm volatile("bar.sync 1, %0;" ::"n"(128 * MH));
      for (int cc = (warp - 2) * 32 + lane; cc < BN; cc += 128 * MH) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 4 * MH; ++w) t += csm[w * BN + cc];
        if (n0 + cc < g.N) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(g.colsum + n0 + cc), "f"(t) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS));
  }
}

// TF32 tensor-pipe peak probe (roofline denominator of the wide path): every CTA keeps one A tile (128 x 32) and one
// B tile (256 x 32) in shared memory (K-major, SWIZZLE_128B: the layout the forward product uses) and one lane issues
// tcgen05.mma.kind::tf32 (M=128, N=256, K=8) back to back into two alternating TMEM accumulators -- no operand
// traffic, so the result is the issue-bound tensor rate at the clocks the part sustains.  2*128*256*8 FLOP per MMA.
__global__ void __launch_bounds__(128) tf32_peak_kernel(int iters, float* out) {
  extern __shared__ unsigned char pk_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pk_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  float* t = reinterpret_cast<float*>(tiles);
  for (int i = threadIdx.x; i < (128 + 256) * 32; i += blockDim.x) t[i] = 1e-3f * (float)((i * 37 + blockIdx.x) % 101 - 50);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile writes -> tensor-core reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc_instr_desc(128, 256, false, false);
      const uint32_t a_addr = smem_u32(tiles), b_addr = a_addr + 128 * 32 * 4;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = tc_smem_desc(a_addr + k * 32, 16, 1024, 2);
          const uint64_t bd = tc_smem_desc(b_addr + k * 32, 16, 1024, 2);
          tc_mma_tf32(tmem_base + (it & 1) * 256, ad, bd, idesc, it > 1 ? 1u : 0u);
        }
      }
      tc_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    float v[32];
    tc_ld32(tmem_base, v);          // warp 0 reads its own lane quadrant (TMEM lanes 0..31)
    if (lane == 0) out[blockIdx.x] = v[0];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

}  // namespace mafb
