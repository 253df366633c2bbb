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

template <int BN>
struct TcCfg {
  static constexpr int STAGES = BN >= 128 ? 3 : 4;
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;     // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
};

template <int BN, bool A_MN, bool B_MN, int EPI, bool RELU>
__global__ void __launch_bounds__(192)
wide_tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ GemmArgs g) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* accum_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
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
  if (warp == 1) {   // TMEM: BN FP32 accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(BN));
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
          for (int a = 0; a < TC_BM / 32; ++a) tma_load_2d(a_dst + a * 4096, &tmA, m0 + 32 * a, k0, &full[s]);
        } else {
          tma_load_2d(a_dst, &tmA, k0, m0, &full[s]);
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
      constexpr uint32_t idesc = tc_instr_desc(TC_BM, BN, A_MN, B_MN);
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
          const uint64_t ad = A_MN ? tc_smem_desc(a_addr + k * 1024, 4096, 512, 1)
                                   : tc_smem_desc(a_addr + k * 32, 16, 1024, 2);
          const uint64_t bd = B_MN ? tc_smem_desc(b_addr + k * 1024, 4096, 512, 1)
                                   : tc_smem_desc(b_addr + k * 32, 16, 1024, 2);
          tc_mma_tf32(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
        }
        tc_commit(&empty[s]);          // smem slot reusable once these MMAs have read it
      }
      tc_commit(accum_full);           // accumulator complete
    }
  } else {
    // ===== epilogue warps: TMEM lane group = warp % 4
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int lg = warp & 3;
    const int m = m0 + lg * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      if (num_kb > 0) tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + c, v);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (m < g.M) {
        const int nb = n0 + c;
        if (EPI == EPI_ATOMIC_MASK) {
          float* crow = g.C + (size_t)m * g.ldc + nb;
          const unsigned char* mrow = g.mask + (size_t)m * g.N + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < g.N && mrow[j]) atomicAdd(crow + j, v[j]);
        } else {
          float* crow = g.C + (size_t)m * g.ldc + nb;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (nb + j >= g.N) break;
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int n = nb + j + q;
              float r = v[j + q];
              if (EPI == EPI_BIAS_ACT) {
                float hh, dd;
                act_eval<RELU>(g.act, r + g.bias[n], hh, dd);
                r = hh;
                if (!RELU && g.aux_out) g.aux_out[(size_t)m * g.ldaux_out + n] = dd;
              } else if (EPI == EPI_BIAS) {
                r = r + g.bias[n];
              } else if (EPI == EPI_MUL_DERIV) {
                const float a = g.aux[(size_t)m * g.ldaux + n];
                r = RELU ? (a > 0.f ? r : 0.f) : r * a;
              } else if (EPI == EPI_ADD) {
                r = r + g.aux[(size_t)m * g.ldaux + n];
              }
              o[q] = r;
            }
            *reinterpret_cast<float4*>(crow + j) = make_float4(o[0], o[1], o[2], o[3]);   // N % 4 == 0
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
  }
}

}  // namespace mafb
