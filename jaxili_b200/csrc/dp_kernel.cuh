// dp_kernel.cuh -- one-shot all-reduce of the flat gradient over NVLink peer memory (pull kernel below, and the
// push exchange inside step_tail_kernel<DP> further down), fused with the
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
constexpr long long kTailWaitCycles = 4000000000LL;   // ~2 s: the blocks of one cooperative launch waiting for each other

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
  long long* trace;                       // debug (nullable): %globaltimer stamps of block 0 in step_tail_kernel
  // push protocol of step_tail_kernel<true>: every rank owns a receive area [2 parities][DP_MAX_WORLD sources]
  // [slot_floats] of 64-bit words {value, epoch} at float offset recv_off of its symmetric buffer
  long long recv_off;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// 64-bit {value, epoch} words: a naturally aligned 64-bit store is single-copy atomic, so a reader that sees the
// expected epoch in the upper half has the value in the lower half -- no fence, no separate flag (the idea of
// NCCL's LL protocol), one NVLink one-way latency instead of store + fence + flag round trips
__device__ __forceinline__ void st_ll(unsigned long long* p, float v, unsigned int epoch) {
  const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_ll(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
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
    const long long patience = *reinterpret_cast<volatile unsigned int*>(a.error) ? 0 : a.timeout_cycles;
    // epochs only grow; signed difference tolerates wrap-around
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {
      if (clock64() - t0 > patience) {
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

// ---------------------------------------------------------------------------------------------------------
// Whole tail of a training step in ONE launch: reduction of the per-CTA partial gradients (reduce_kernel),
// [DP: one-shot exchange with the peers, as dp_allreduce_kernel], global norm, clip + schedule + Adam +
// re-pack of the weight images (clip_adam_kernel).  The global norm needs every block's partial: each block
// publishes its partial as a 64-bit {value, launch epoch} word and polls everybody else's (one L2 round trip
// instead of a grid barrier and a re-read); the kernel is launched cooperatively (all ceil(P/128) blocks
// co-resident: narrow models only) and each parameter's gradient stays in the register of the thread that
// produced it across the wait.
template <bool DP>
__global__ void __launch_bounds__(RED_THREADS, 2)   // two blocks per SM: all ceil(P/128) blocks co-resident
step_tail_kernel(const __grid_constant__ Plan P, const __grid_constant__ OptArgs O, const float* __restrict__ partial,
                 int G, const unsigned char* __restrict__ mask, float* __restrict__ grad,
                 const float* __restrict__ partial_loss, float* __restrict__ loss, float* __restrict__ normpart,
                 float* __restrict__ params, float* __restrict__ m, float* __restrict__ v, int* __restrict__ step,
                 float* __restrict__ wimg, unsigned long long* __restrict__ nw,
                 const __grid_constant__ DpArgs a) {
  __shared__ __align__(16) float red[RED_PARTS][RED_IDX];
  __shared__ float wsum[RED_IDX / 32];
  __shared__ float lsum[RED_IDX / 32];
  __shared__ float s_norm;
  __shared__ int s_ok, s_fail;
  const int i = threadIdx.x % RED_IDX, p = threadIdx.x / RED_IDX;
  const int idx = blockIdx.x * RED_IDX + i;
  auto tstamp = [&](int k) {
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (k == 0) a.trace[7] = (long long)now - a.trace[6];   // previous tail's end -> this tail's start
      a.trace[k] = (long long)now;
    }
  };
  tstamp(0);
  // operands of the update: independent of the flow kernel, fetched before the programmatic-dependency wait
  float pv = 0.f, mv = 0.f, vv = 0.f;
  unsigned char mk = 0;
  if (p == 0 && idx < P.P) {
    pv = params[idx];
    mk = mask[idx];
    if (O.kind != 1) { mv = m[idx]; vv = v[idx]; }
  }
  griddep_launch_dependents();     // the next step's flow kernel may start launching (it waits for this grid's end)
  griddep_wait();                  // the flow kernel's partial gradients and step counter are complete
  const int t = step[0];
  // launch epoch of the norm words (nw[0] = launches so far, bumped by block 0 once it has seen every word)
  const unsigned int lep = ((unsigned int)__ldcg(nw) + 1u) & 0x7fffffffu;   // 31 bits: bit 31 of a word's tag is the give-up flag
  unsigned long long* const words = nw + 1;
  auto publish = [&](float tt, bool ok) {
    // {squared-norm partial, launch epoch | bit 31: this block gave up on a peer}: the value travels with its tag
    const unsigned int tag = lep | (ok ? 0u : 0x80000000u);
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(tt);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(words + blockIdx.x), "l"(w) : "memory");
  };
  // DP: the locally reduced gradient is PUSHED into every rank's receive area of the coming epoch
  // (= count + 1; parity alternates): remote stores are fire-and-forget, the later sum reads local memory only
  const unsigned int epoch = (unsigned int)t + 1u;
  // index (in 64-bit words) of this rank's lane in a receive area of the epoch's parity
  const long long roff = DP ? ((long long)(epoch & 1u) * DP_MAX_WORLD + a.rank) * a.slot_floats : 0;
  reduce_partials(partial, P.pgs, G, P.P, red);
  __syncthreads();
  float g = 0.f;
  if (p == 0) {
    if (idx < P.P) {
      g = red_parts_sum(red, i);
      g = mk ? g : 0.f;
      if (DP) {
#pragma unroll
        for (int r = 0; r < DP_MAX_WORLD; ++r)
          if (r < a.world)
            st_ll(reinterpret_cast<unsigned long long*>(const_cast<float*>(a.sym[r]) + a.recv_off) + roff + idx, g, epoch);
      } else {
        grad[idx] = g;
      }
    }
    if (!DP) {
      const float sq = warp_sum(g * g);
      if ((i & 31) == 0) wsum[i >> 5] = sq;
    }
  } else if (blockIdx.x == 0 && p == 1) {   // second part-group of block 0 reduces the loss
    float l = 0.f;
    for (int q = i; q < G; q += RED_IDX) l += partial_loss[q];
    l = warp_sum(l);
    if ((i & 31) == 0) lsum[i >> 5] = l;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!DP) {
      float tt = 0.f;
      for (int w = 0; w < RED_IDX / 32; ++w) tt += wsum[w];
      publish(tt, true);
    }
    if (blockIdx.x == 0) {
      float l = 0.f;
      for (int w = 0; w < RED_IDX / 32; ++w) l += lsum[w];
      if (DP) {
        for (int r = 0; r < a.world; ++r)
          st_ll(reinterpret_cast<unsigned long long*>(const_cast<float*>(a.sym[r]) + a.recv_off) + roff + ((P.P + 3) & ~3),
                l, epoch);
      } else {
        loss[0] = l;
      }
    }
    s_ok = 1;
    s_fail = 0;
    tstamp(1);
  }
  tstamp(2);
  __syncthreads();
  if (p != 0) return;
  if (DP) {
    // every value arrives together with its epoch: spin on the 64-bit word itself (local memory)
    const unsigned long long* mine =
        reinterpret_cast<const unsigned long long*>(a.sym[a.rank] + a.recv_off) +
        (long long)(epoch & 1u) * DP_MAX_WORLD * a.slot_floats;
    const long long t0 = clock64();
    // once a peer has been declared missing every later step gives up at once (no 20 s spin per step)
    const long long patience = *reinterpret_cast<volatile unsigned int*>(a.error) ? 0 : a.timeout_cycles;
    auto take = [&](long long word) -> float {
      unsigned long long w = ld_ll(mine + word);
      while ((unsigned int)(w >> 32) != epoch) {
        if (clock64() - t0 > patience) {
          s_ok = 0;
          *a.error = 1u;
          break;
        }
        w = ld_ll(mine + word);
      }
      return __uint_as_float((unsigned int)w);
    };
    tstamp(3);
    g = 0.f;
    if (idx < P.P) {
#pragma unroll
      for (int r = 0; r < DP_MAX_WORLD; ++r)                   // fixed order: replicas stay bit-identical
        if (r < a.world) g += take((long long)r * a.slot_floats + idx);
      grad[idx] = g;
    }
    float l_tot = 0.f;
    if (blockIdx.x == 0 && i == 0)
      for (int r = 0; r < a.world; ++r) l_tot += take((long long)r * a.slot_floats + ((P.P + 3) & ~3));
    const float sq = warp_sum(g * g);
    if ((i & 31) == 0) wsum[i >> 5] = sq;
    asm volatile("bar.sync 1, %0;" ::"n"(RED_IDX));
    if (i == 0) {
      float tt = 0.f;
      for (int w = 0; w < RED_IDX / 32; ++w) tt += wsum[w];
      publish(tt, s_ok != 0);
      if (blockIdx.x == 0) loss[0] = s_ok ? l_tot : __int_as_float(0x7fc00000);
      tstamp(4);
    }
  }
  // what the update needs besides the norm: schedule, bias corrections, the parameter's slots in the weight images
  const float lr_t = lr_schedule(O, t);
  const float tn = (float)(t + 1);
  const float d1 = 1.f - powf(O.b1, tn), d2 = 1.f - powf(O.b2, tn);
  long long o1 = 0, o2 = -1;
  if (idx < P.P) image_offsets(P, idx, o1, o2);
  // Global norm without a grid barrier: every block polls every block's word (both of a thread's words requested
  // together) and sums them in block order -> identical scale everywhere, one L2 round trip behind the last block.
  // A peer went missing (time-out above, in any block of this rank): the block's tag says so, every block reads every
  // tag, and the whole rank leaves params / m / v / step / the weight images untouched -- a stale word must never
  // reach the optimiser state.  The loss is NaN and mafb200_dp_error reports it to the host.
  float ns = 0.f;
  {
    const long long tw = clock64();
    const long long patience = DP ? a.timeout_cycles : kTailWaitCycles;
    unsigned int bad = DP ? (__ldcg(a.error) != 0u ? 1u : 0u) : 0u;     // sticky: declared missing in an earlier step
    auto ldw = [&](int q) {
      unsigned long long w;
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(words + q) : "memory");
      return w;
    };
    auto settle = [&](int q, unsigned long long w) -> float {
      while (((unsigned int)(w >> 32) & 0x7fffffffu) != lep) {
        if (clock64() - tw > patience) { w = 1ull << 63; break; }
        w = ldw(q);
      }
      bad |= (unsigned int)(w >> 63);
      return __uint_as_float((unsigned int)w);
    };
    const int nblk = (int)gridDim.x;
    for (int q = i; q < nblk; q += 2 * RED_IDX) {
      const bool two = q + RED_IDX < nblk;
      const unsigned long long w0 = ldw(q), w1 = two ? ldw(q + RED_IDX) : 0ull;
      ns += settle(q, w0);
      if (two) ns += settle(q + RED_IDX, w1);
    }
    if (bad) atomicOr(&s_fail, 1);
  }
  ns = warp_sum(ns);
  if ((i & 31) == 0) wsum[i >> 5] = ns;
  asm volatile("bar.sync 1, %0;" ::"n"(RED_IDX));
  if (i == 0) {
    float tot = 0.f;
    for (int w = 0; w < RED_IDX / 32; ++w) tot += wsum[w];
    s_norm = sqrtf(tot);
  }
  asm volatile("bar.sync 1, %0;" ::"n"(RED_IDX));
  tstamp(5);
  const bool dp_fail = s_fail != 0;
  if (dp_fail && blockIdx.x == 0 && i == 0) {
    loss[0] = __int_as_float(0x7fc00000);
    if (DP) *a.error = 1u;
  }
  const float gnorm = s_norm;
  if (idx < P.P && !dp_fail) {
    if (!(gnorm < O.clip)) g = (g / gnorm) * O.clip;
    float upd;
    if (O.kind == 1) {
      upd = g + O.weight_decay * pv;
    } else {
      const float mm = O.b1 * mv + (1.f - O.b1) * g;
      const float v2 = O.b2 * vv + (1.f - O.b2) * g * g;
      m[idx] = mm;
      v[idx] = v2;
      const float mhat = mm / d1;
      const float vhat = v2 / d2;
      upd = mhat / (sqrtf(vhat) + O.eps);
      if (O.kind == 2) upd += O.weight_decay * pv;
    }
    pv = pv - lr_t * upd;
    params[idx] = pv;
    const float iv = mk ? pv : 0.f;
    wimg[o1] = iv;
    if (o2 >= 0) wimg[o2] = iv;
  }
  // block 0 has seen every block's word: every block has read step[0] and nw[0]
  if (blockIdx.x == 0 && i == 0) *nw = (unsigned long long)lep;
  if (blockIdx.x == 0 && i == 0 && !dp_fail) step[0] = t + 1;
  tstamp(6);
#ifdef MAFB_CTA_TIMES
  if (a.trace && threadIdx.x == 0) {   // debug: every block's exit time (trace + 900 is the base: slots 600..)
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    a.trace[-300 + (int)blockIdx.x] = (long long)now;
  }
#endif
}

}  // namespace mafb
