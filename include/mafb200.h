/*
 * mafb200.h -- C ABI of the B200-native ConditionalMAF hot path.
 *
 * This is the drop-in boundary for jaxili's data-parallel hot path.  The reference
 * (sachaguer/jaxili) has no FFI seam of its own (it is pure Python on JAX); the entry
 * points below are what an XLA-FFI / ctypes binding for that path binds, one per
 * reference call site (paths relative to the jaxili repository root):
 *
 *   mafb200_log_prob          <- model.apply(..., method="log_prob")
 *                                jaxili/model.py:1073-1098 -> 903-921 -> 848-874 -> 718-743
 *                                (callers: jaxili/loss.py:57,83,
 *                                 jaxili/posterior/direct_posterior.py:106,
 *                                 jaxili/train.py:337-340 eval_step)
 *   mafb200_loss_grad         <- jax.value_and_grad(loss_nll_npe / loss_nll_nle)
 *                                jaxili/train.py:330-332, jaxili/loss.py:35-83
 *   mafb200_log_prob_grad_y   <- jax.grad of MCMCPosterior.log_likelihood w.r.t. theta (NLE)
 *                                jaxili/posterior/mcmc_posterior.py:139-159, 194-201
 *   mafb200_hmc_trajectory    <- numpyro HMC leapfrog integration inside MCMCPosterior.sample
 *                                jaxili/posterior/mcmc_posterior.py:78-118, 248-305
 *   mafb200_pipeline_*        <- TrainerModule.train_epoch's host loop             jaxili/train.py:427-451
 *   mafb200_train_step        <- TrainerModule.create_functions().train_step    jaxili/train.py:330-335
 *   mafb200_clip_adam         <- state.apply_gradients(grads) with the optax chain built in
 *                                jaxili/train.py:246-300 (clip_by_global_norm -> adam(warmup_cosine))
 *   mafb200_forward           <- ConditionalMAF.__call__ (u, log_det)  jaxili/model.py:848-874
 *   mafb200_inverse           <- ConditionalMAF.backward (x, log_det)  jaxili/model.py:876-901
 *   mafb200_sample_from_noise <- model.apply(..., method="sample") with the base noise given
 *   mafb200_sample_philox        jaxili/model.py:1116-1126 -> 923-947 -> 876-901 -> 745-773
 *                                (caller: jaxili/posterior/direct_posterior.py:70-72)
 *   mafb200_create            <- ConditionalMAF.setup / ConditionalMADE._create_masks
 *                                jaxili/model.py:829-846, 581-663 (degrees in, masks derived)
 *                                + NDE_w_Standardization constants, jaxili/inference/npe.py:383-399
 *
 * Conventions
 *   - plain C types only; every device pointer is a caller-owned CUDA device pointer to
 *     FP32 row-major data, 16-byte aligned; the library owns only the opaque handle and
 *     its private scratch.  Nothing allocates on the hot path.
 *   - every compute entry enqueues on the caller's stream (a cudaStream_t passed as
 *     void*) and returns without synchronising; all are CUDA-graph capturable.
 *   - return value 0 = ok, non-zero = error; the message is in mafb200_last_error()
 *     (thread-local).  There is no CPU fallback: without a CUDA device every compute
 *     entry fails.
 *   - flat parameter order: for MAF layer k = 0..L-1, for masked linear i: kernel_i
 *     (in_i x out_i, row-major, flax Dense convention) then bias_i (out_i).
 */
#ifndef MAFB200_H_
#define MAFB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAFB200_MAX_HIDDEN 5

/* activation enum: names follow jax.nn (jaxili/inventory/func_dict.py:28-31) */
enum {
  MAFB200_ACT_RELU = 0,
  MAFB200_ACT_SILU = 1,
  MAFB200_ACT_TANH = 2,
  MAFB200_ACT_GELU = 3, /* jax.nn.gelu default: tanh approximation */
  MAFB200_ACT_ELU = 4,
  MAFB200_ACT_SIGMOID = 5,
  MAFB200_ACT_SOFTPLUS = 6,
  MAFB200_ACT_LEAKY_RELU = 7 /* negative_slope 0.01 */
};

/* flags for the compute entries */
enum {
  /* the handle's packed (mask-folded) weight image already reflects `params`:
     skip the re-pack launch.  clip_adam leaves the image current. */
  MAFB200_PACKED_CURRENT = 1,
  /* clip_adam: the handle's squared-norm partials already describe `grad` (true right
     after loss_grad on the same handle when grad was not modified, e.g. no all-reduce):
     skip the norm launch. */
  MAFB200_NORM_CURRENT = 2,
  /* loss_grad: write the reduced gradient and loss into this rank's symmetric (peer-mapped) slot
     instead of out_grad / out_loss; mafb200_dp_allreduce then sums the ranks' slots.
     train_step / pipeline_create: data-parallel step, the exchange runs inside the tail kernel. */
  MAFB200_DP_SLOT = 4,
  /* loss_grad, measurement hook: launch the fused flow kernel only and leave out_loss / out_grad untouched
     (the per-CTA partial gradients stay in the handle).  bench.py times back-to-back launches of it for the
     roofline figure. */
  MAFB200_FLOW_ONLY = 8
};

typedef struct mafb200_handle_s* mafb200_handle;

/* ConditionalMAF(n_in, n_cond, n_layers, layers, activation, use_reverse)
   jaxili/model.py:821-827 */
typedef struct {
  int32_t n_in;
  int32_t n_cond;
  int32_t n_layers;
  int32_t n_hidden;                    /* len(layers) >= 1 */
  int32_t hidden[MAFB200_MAX_HIDDEN];  /* layers */
  int32_t activation;                  /* MAFB200_ACT_* */
  int32_t use_reverse;
} MafDesc;

/* optimizer_hparams of TrainerModule.init_optimizer, jaxili/train.py:257-292 */
typedef struct {
  float lr;            /* peak learning rate */
  float warmup;        /* warmup_steps (NPE.train passes 0.1) */
  float decay_steps;   /* cosine horizon, same unit as warmup */
  float init_factor;   /* init_value = init_factor*lr  (0.01) */
  float end_factor;    /* end_value  = end_factor*lr   (0.01) */
  float b1, b2, eps;   /* adam: 0.9, 0.999, 1e-8 */
  float clip;          /* clip_by_global_norm max_norm (5.0) */
  float weight_decay;  /* adamw only */
  int32_t kind;        /* 0 adam, 1 sgd (+decayed weights), 2 adamw */
} MafOptDesc;

/* State of n vectorised HMC chains over  theta -> log_prior(theta) + log p(x | theta)  in the unconstrained
   space z of the prior's support (what numpyro's HMC/NUTS do inside MCMCPosterior,
   jaxili/posterior/mcmc_posterior.py:179-305).  All pointers are DEVICE pointers owned by the caller.
   prior_kind[c]: 0 = Normal(loc = lo[c], scale = hi[c]) on the real line (theta = z),
                  1 = Uniform(lo[c], hi[c]) (theta = lo + (hi - lo) sigmoid(z)). */
typedef struct {
  const float* eps;            /* [1]  leapfrog step size (device scalar, so adaptation needs no re-capture) */
  const float* inv_mass;       /* [C]  diagonal inverse mass matrix */
  const int32_t* prior_kind;   /* [C] */
  const float* lo;             /* [C] */
  const float* hi;             /* [C] */
  float* z;                    /* [n][C] in/out  position */
  float* p;                    /* [n][C] in/out  momentum */
  float* g;                    /* [n][C] in/out  gradient of the log density at z */
  float* logdens;              /* [n]    out     log density at z (up to a constant) */
  float* theta;                /* [n][C] out     constrained position */
  float* scratch;              /* [n][C + 1]     work space */
} MafHmcState;

const char* mafb200_last_error(void);
int mafb200_version(void);

/* degrees: for each MAF layer k, the concatenated hidden degree vectors m^1..m^n
   (sum(hidden) int32 each), exactly as np.random.randint produced them
   (jaxili/model.py:613-616); masks are derived as in 621-657.
   shift/scale/mean/std: host arrays (n_in, n_in, n_cond, n_cond) of the
   ScalarAffine / Standardizer; NULL = identity. */
int mafb200_create(const MafDesc* desc, const int32_t* degrees,
                   const float* shift, const float* scale,
                   const float* mean, const float* std,
                   int device, mafb200_handle* out);
int mafb200_destroy(mafb200_handle h);
int mafb200_set_standardization(mafb200_handle h, const float* shift, const float* scale,
                                const float* mean, const float* std);

/* debug: phase-boundary clock stamps of CTA 0 of the flow kernel go to this device buffer of
   1001 int64 ((id, clock) pairs, count at [1000]); NULL switches tracing off (default). */
int mafb200_debug_trace(mafb200_handle h, long long* trace_dev);

/* measurement aid: bracket the fused flow kernel (alone) of every following log_prob / forward /
   loss_grad call with CUDA events on the launch stream; mafb200_flow_elapsed returns the summed
   device time and the number of bracketed launches since timing was switched on.  Not for use
   under CUDA-graph capture. */
int mafb200_flow_timing(mafb200_handle h, int32_t enable);
int mafb200_flow_elapsed(mafb200_handle h, double* total_ms, int64_t* launches);

/* Data parallel over NVLink peer memory (one process per GPU, at most 8 ranks of one node).
   dp_init allocates this rank's symmetric buffer and returns its 64-byte CUDA IPC handle; the host
   exchanges handles (e.g. torch.distributed.all_gather) and passes all of them, rank-major, to
   dp_connect; dp_bind_step names the device step counter (the exchange epoch and the slot parity are
   derived from it on the device, so the whole step is CUDA-graph capturable).
   Per step, either  train_step(..., MAFB200_DP_SLOT)  -- two launches; the tail kernel pushes every reduced
   gradient entry to all peers as a 64-bit {value, epoch} word and sums the world's words locally --
   or the three-call form  loss_grad(..., MAFB200_DP_SLOT) -> dp_allreduce -> clip_adam(..., MAFB200_NORM_CURRENT),
   where dp_allreduce is a one-shot pull all-reduce (every rank reads every peer's slot).  Both sum in a
   fixed rank order, so replicas stay bit-identical, and replace ncclAllReduce for this latency-bound
   message.  Every rank must make the same calls the same number of times.  dp_reset (between host
   barriers) forgets all epochs: required whenever the step counter is replaced or rewound.  dp_error
   reports a peer time-out (synchronises). */
int mafb200_dp_init(mafb200_handle h, int32_t rank, int32_t world, void* ipc_handle_out);
int mafb200_dp_connect(mafb200_handle h, const void* all_handles);
int mafb200_dp_bind_step(mafb200_handle h, const int32_t* step);   /* device optax count = exchange epoch - 1 */
int mafb200_dp_reset(mafb200_handle h);   /* all ranks, between host barriers, when the count restarts */
int mafb200_dp_allreduce(mafb200_handle h, float* grad_out, float* loss_out, void* stream);
int mafb200_dp_error(mafb200_handle h, int32_t* out);

/* Models whose MADE does not fit shared memory run layer by layer ("wide" path).  Its GEMM core is
   tcgen05 TF32 (core = 1, default; FP32 accumulate, operands truncated to TF32: log_prob agrees
   with FP32 to ~1e-3 absolute) or FP32 FFMA (core = 0, same 1e-5 tolerance as the
   fused narrow kernel).  No effect on models that use the fused kernel. */
int mafb200_set_wide_core(mafb200_handle h, int32_t core);

/* Two fused kernels serve the narrow path (csrc/flow_kernel.cuh: barrier-phased, csrc/chain_kernel.cuh: row-owning
   warp pairs + asynchronous weight-gradient warps); same results to rounding.  mode 0 (default): chain kernel for
   the training pass, barrier-phased kernel for forward-only launches; 1: barrier-phased always; 2: chain kernel
   wherever it applies.  Also settable per process with MAFB200_FLOW=auto|phased|chain. */
int mafb200_set_flow_kernel(mafb200_handle h, int32_t mode);

/* test hook: one GEMM of the wide path (see mafb200.cu); not part of the reference-facing surface */
int mafb200_debug_gemm(int32_t variant, int32_t core, const float* A, const float* B, float* C,
                       int32_t M, int32_t N, int32_t K, const float* zeros_n, const float* zeros_mn,
                       const uint8_t* mask, void* stream);

int64_t mafb200_param_count(mafb200_handle h);
/* writes the 0/1 mask in flat parameter order (biases 1) to a host buffer of param_count bytes */
int mafb200_get_mask(mafb200_handle h, uint8_t* mask_host);
/* bytes of dynamic shared memory, rows per tile and CTAs the flow kernel will use (for reports) */
int mafb200_query(mafb200_handle h, int32_t* smem_bytes, int32_t* tile_rows,
                  int32_t* max_ctas, int32_t* stash_all);

/* mask-fold + pad + transpose `params` into the handle's packed image */
int mafb200_pack(mafb200_handle h, const float* params, void* stream);

/* out_logp[b] = log p(x_b | y_b), b < B.  x: B x n_in, y: B x n_cond. */
int mafb200_log_prob(mafb200_handle h, const float* params, const float* x, const float* y,
                     int64_t B, float* out_logp, int32_t flags, void* stream);

/* ConditionalMAF.__call__ (jaxili/model.py:848-874): any of out_logp [B], out_u [B x n_in]
   (base-space point after the last flip) and out_logdet [B] (summed MAF log|det J|, without
   the standardisation constant) may be NULL. */
int mafb200_forward(mafb200_handle h, const float* params, const float* x, const float* y,
                    int64_t B, float* out_logp, float* out_u, float* out_logdet,
                    int32_t flags, void* stream);

/* eval_step + eval_model's accumulation (jaxili/train.py:337-340, 453-487) on the device:
   acc[0] += -sum_b log_prob(x_b | y_b), acc[1] += B (acc: 2 device floats, zeroed by the caller before the loop).
   The size-weighted mean loss of a loader is acc[0] / acc[1], read once after the loop. */
int mafb200_eval_accumulate(mafb200_handle h, const float* params, const float* x, const float* y,
                            int64_t B, float* acc, int32_t flags, void* stream);

/* out_loss[0] = -inv_global_B * sum_b log p(x_b|y_b); out_grad = d out_loss / d params
   (P floats, masked entries exactly 0).  If batch_index != NULL the batch is rows
   [(batch_index[0] % n_batches) * B, +B) of x / y (device-side batch selection for
   graph replay); otherwise rows [0, B). */
int mafb200_loss_grad(mafb200_handle h, const float* params, const float* x, const float* y,
                      int64_t B, float inv_global_B,
                      const int32_t* batch_index, int64_t n_batches,
                      float* out_loss, float* out_grad, int32_t flags, void* stream);

/* out_logp[b] (nullable) = log p(x_b | y_b), out_dy[b][c] = d log p(x_b | y_b) / d y_b[c]: gradient with
   respect to the conditioning input -- what NUTS/HMC over an NLE likelihood differentiates
   (jaxili/posterior/mcmc_posterior.py:139-159, 194-201).  Fused (narrow) path only. */
int mafb200_log_prob_grad_y(mafb200_handle h, const float* params, const float* x, const float* y,
                            int64_t B, float* out_logp, float* out_dy, int32_t flags, void* stream);

/* n_steps leapfrog steps of all n chains for the observation rows x [n][n_in] (the observation repeated per
   chain): 3 launches per step (drift -> fused flow forward+reverse -> kick), enqueued without synchronising.
   n_steps = 0 evaluates theta, g and logdens at the current z without moving (initialisation).
   Replaces the potential-energy/gradient evaluations of numpyro inside MCMCPosterior._hmc_numpyro /
   _nuts_numpyro, jaxili/posterior/mcmc_posterior.py:194-305. */
int mafb200_hmc_trajectory(mafb200_handle h, const float* params, const float* x, int64_t n,
                           const MafHmcState* state, int32_t n_steps, int32_t flags, void* stream);

/* The whole single-rank step in one call: train_step of jaxili/train.py:330-335 (value_and_grad of the NLL +
   state.apply_gradients).  Same arguments as loss_grad followed by clip_adam; out_loss / out_grad receive the
   step's loss and (unclipped) gradient, params / m / v / step are updated in place.  On narrow models the
   tail (gradient reduction, norm, clip, Adam, re-pack) is a single cooperative kernel. */
int mafb200_train_step(mafb200_handle h, float* params, const float* x, const float* y, int64_t B,
                       float inv_global_B, const int32_t* batch_index, int64_t n_batches, float* out_loss,
                       float* out_grad, float* m, float* v, int32_t* step, const MafOptDesc* opt, int32_t flags,
                       void* stream);

/* Host-batch pipeline: TrainerModule.train_epoch's loop over host batches (jaxili/train.py:427-451, which
   relies on JAX's asynchronous dispatch and reads the loss once per epoch) as a native executor.  Batch i+1
   is copied host->device on a copy stream while step i (one CUDA graph: mafb200_train_step, or loss_grad +
   dp_allreduce + clip_adam with MAFB200_DP_SLOT) runs; every step's loss is copied device->host into a
   pinned ring on a third stream.  params / m / v / step / grad are the caller's device buffers (as in
   mafb200_train_step); after_stream orders the pipeline behind the caller's earlier work.
   create runs one throw-away step on a saved copy of the optimiser state (restored before it returns);
   capture builds the graphs; submit enqueues one step and returns; drain blocks until all submitted steps
   are done and returns their losses. */
typedef struct mafb200_pipeline_s* mafb200_pipeline;
int mafb200_pipeline_create(mafb200_handle h, int64_t rows, float* params, float* m, float* v, int32_t* step,
                            float* grad, const MafOptDesc* opt, float inv_global_B, int32_t flags,
                            void* after_stream, mafb200_pipeline* out);
int mafb200_pipeline_capture(mafb200_pipeline p);
int mafb200_pipeline_fence(mafb200_pipeline p, void* after_stream);
int mafb200_pipeline_submit(mafb200_pipeline p, const float* x, const float* y, int32_t pinned);
int mafb200_pipeline_drain(mafb200_pipeline p, float* out_losses, int64_t cap, int64_t* n, float* last_loss_dev);
int mafb200_pipeline_destroy(mafb200_pipeline p);

/* optax.chain(clip_by_global_norm, adam(schedule)) + apply_gradients, in place on
   params / m / v; step[0] (device int32) is the optax count and is incremented.
   Leaves the packed image current. */
int mafb200_clip_adam(mafb200_handle h, float* params, const float* grad, float* m, float* v,
                      int32_t* step, const MafOptDesc* opt, int32_t flags, void* stream);

/* theta[r] = scale * MAF^-1(u[r] | (y_obs[r / rows_per_obs] - mean)/std) + shift.
   u, out: N x n_in; y_obs: n_obs x n_cond with N <= n_obs * rows_per_obs. */
int mafb200_sample_from_noise(mafb200_handle h, const float* params, const float* u,
                              const float* y_obs, int64_t n_obs, int64_t rows_per_obs,
                              int64_t N, float* out, int32_t flags, void* stream);
/* ConditionalMAF.backward (jaxili/model.py:876-901): per-row conditioning y [N x n_cond];
   out_x [N x n_in], out_logdet [N] = sum over layers of sum_d min(-logp_d/2, 10) (nullable). */
int mafb200_inverse(mafb200_handle h, const float* params, const float* u, const float* y,
                    int64_t N, float* out_x, float* out_logdet, int32_t flags, void* stream);
/* same with u ~ N(0,I) drawn in-kernel: Philox4x32-10, subsequence = row_offset + r */
int mafb200_sample_philox(mafb200_handle h, const float* params, uint64_t seed,
                          uint64_t row_offset, const float* y_obs, int64_t n_obs,
                          int64_t rows_per_obs, int64_t N, float* out, int32_t flags,
                          void* stream);

/* FP32 FFMA micro-benchmark used as the roofline denominator (bench.py): runs `iters`
   dependent-chain-free FFMA blocks on every SM, returns achieved TFLOP/s. */
int mafb200_ffma_peak(int device, int32_t iters, float* out_tflops);
/* the same with packed FFMA2 (packed = 1: fma.rn.f32x2, what the fused kernels issue) or plain FFMA (packed = 0) */
int mafb200_fp32_peak(int device, int32_t iters, int32_t packed, float* out_tflops);
/* TF32 tcgen05 tensor-pipe peak (M=128, N=256, K=8 MMAs issued back to back on resident shared-memory operands):
   roofline denominator of the wide path; `iters` k-blocks of 4 MMAs per CTA, one CTA per SM. */
int mafb200_tf32_peak(int device, int32_t iters, float* out_tflops);

#ifdef __cplusplus
}
#endif
#endif /* MAFB200_H_ */
