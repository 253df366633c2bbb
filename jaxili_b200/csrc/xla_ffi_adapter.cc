// xla_ffi_adapter.cc -- XLA FFI (jax.ffi) handlers wrapping the C ABI of include/mafb200.h.
//
// NOT compiled in this image: JAX / xla/ffi/api/ffi.h are not installed.  A jaxili
// maintainer builds this file against their jaxlib headers:
//
//   g++ -shared -fPIC -O2 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       xla_ffi_adapter.cc -L. -lmafb200 -o libmafb200_ffi.so
//
// and registers the handlers as shown in INTEGRATION.md.  The MAF handle is created once on the
// Python side (ctypes -> mafb200_create) and passed as an int64 attribute.
#if __has_include("xla/ffi/api/ffi.h")
#include "xla/ffi/api/ffi.h"

#include "../../include/mafb200.h"

namespace ffi = xla::ffi;

static ffi::Error Check(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, mafb200_last_error());
}

// log_prob(params[P], x[B,D], y[B,C]) -> logp[B]      (jaxili/model.py:1073-1098, 903-921)
static ffi::Error MafLogProbImpl(cudaStream_t stream, int64_t handle, ffi::Buffer<ffi::F32> params,
                                 ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> y,
                                 ffi::ResultBuffer<ffi::F32> logp) {
  auto h = reinterpret_cast<mafb200_handle>(handle);
  return Check(mafb200_log_prob(h, params.typed_data(), x.typed_data(), y.typed_data(),
                                x.dimensions()[0], logp->typed_data(), 0, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MafLogProb, MafLogProbImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// loss_grad(params, x, y) -> (loss[1], grad[P])        (jaxili/train.py:330-332)
static ffi::Error MafLossGradImpl(cudaStream_t stream, int64_t handle, float inv_global_b,
                                  ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> x,
                                  ffi::Buffer<ffi::F32> y, ffi::ResultBuffer<ffi::F32> loss,
                                  ffi::ResultBuffer<ffi::F32> grad) {
  auto h = reinterpret_cast<mafb200_handle>(handle);
  return Check(mafb200_loss_grad(h, params.typed_data(), x.typed_data(), y.typed_data(),
                                 x.dimensions()[0], inv_global_b, nullptr, 1, loss->typed_data(),
                                 grad->typed_data(), 0, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MafLossGrad, MafLossGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<float>("inv_global_b")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// sample_from_noise(params, u[N,D], y_obs[n_obs,C]) -> theta[N,D]   (jaxili/model.py:1116-1126)
static ffi::Error MafSampleImpl(cudaStream_t stream, int64_t handle, int64_t rows_per_obs,
                                ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> u,
                                ffi::Buffer<ffi::F32> y_obs, ffi::ResultBuffer<ffi::F32> out) {
  auto h = reinterpret_cast<mafb200_handle>(handle);
  return Check(mafb200_sample_from_noise(h, params.typed_data(), u.typed_data(), y_obs.typed_data(),
                                         y_obs.dimensions()[0], rows_per_obs, u.dimensions()[0],
                                         out->typed_data(), 0, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MafSample, MafSampleImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<int64_t>("rows_per_obs")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// optimiser hyper-parameters arrive as scalar attributes (TrainerModule.init_optimizer, jaxili/train.py:257-292)
static MafOptDesc OptFrom(float lr, float warmup, float decay_steps, float clip, float weight_decay, int32_t kind) {
  MafOptDesc o{};
  o.lr = lr; o.warmup = warmup; o.decay_steps = decay_steps;
  o.init_factor = 0.01f; o.end_factor = 0.01f;
  o.b1 = 0.9f; o.b2 = 0.999f; o.eps = 1e-8f;
  o.clip = clip; o.weight_decay = weight_decay; o.kind = kind;
  return o;
}

// clip_adam(params, grad, m, v, step) -> (params', m', v', step')   optax.chain(clip_by_global_norm, adam(schedule))
// + apply_gradients (jaxili/train.py:246-300, 333).  The C ABI updates in place: the outputs alias the inputs
// (input_output_aliases={0: 0, 2: 1, 3: 2, 4: 3} on the jax.ffi.ffi_call side), XLA then hands the same buffers in.
static ffi::Error MafClipAdamImpl(cudaStream_t stream, int64_t handle, float lr, float warmup, float decay_steps,
                                  float clip, float weight_decay, int32_t kind, ffi::Buffer<ffi::F32> params,
                                  ffi::Buffer<ffi::F32> grad, ffi::Buffer<ffi::F32> m, ffi::Buffer<ffi::F32> v,
                                  ffi::Buffer<ffi::S32> step, ffi::ResultBuffer<ffi::F32> params_out,
                                  ffi::ResultBuffer<ffi::F32> m_out, ffi::ResultBuffer<ffi::F32> v_out,
                                  ffi::ResultBuffer<ffi::S32> step_out) {
  auto h = reinterpret_cast<mafb200_handle>(handle);
  if (params_out->typed_data() != params.typed_data() || m_out->typed_data() != m.typed_data() ||
      v_out->typed_data() != v.typed_data() || step_out->typed_data() != step.typed_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "MafClipAdam updates in place: alias outputs to inputs");
  const MafOptDesc o = OptFrom(lr, warmup, decay_steps, clip, weight_decay, kind);
  return Check(mafb200_clip_adam(h, params_out->typed_data(), grad.typed_data(), m_out->typed_data(),
                                 v_out->typed_data(), step_out->typed_data(), &o, 0, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MafClipAdam, MafClipAdamImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<float>("lr").Attr<float>("warmup").Attr<float>("decay_steps")
                                  .Attr<float>("clip").Attr<float>("weight_decay").Attr<int32_t>("kind")
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::S32>>());

// train_step(params, m, v, step, x, y) -> (params', m', v', step', loss[1])   jaxili/train.py:330-335 in one call:
// fused forward + reverse pass, then the cooperative tail kernel (reduce, clip, schedule, Adam, re-pack).
// `grad` is a scratch output of P floats (the clipped-before gradient, as mafb200_train_step leaves it).
static ffi::Error MafTrainStepImpl(cudaStream_t stream, int64_t handle, float inv_global_b, float lr, float warmup,
                                   float decay_steps, float clip, float weight_decay, int32_t kind,
                                   ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> m, ffi::Buffer<ffi::F32> v,
                                   ffi::Buffer<ffi::S32> step, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> y,
                                   ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::F32> m_out,
                                   ffi::ResultBuffer<ffi::F32> v_out, ffi::ResultBuffer<ffi::S32> step_out,
                                   ffi::ResultBuffer<ffi::F32> loss, ffi::ResultBuffer<ffi::F32> grad) {
  auto h = reinterpret_cast<mafb200_handle>(handle);
  if (params_out->typed_data() != params.typed_data() || m_out->typed_data() != m.typed_data() ||
      v_out->typed_data() != v.typed_data() || step_out->typed_data() != step.typed_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "MafTrainStep updates in place: alias outputs to inputs");
  const MafOptDesc o = OptFrom(lr, warmup, decay_steps, clip, weight_decay, kind);
  return Check(mafb200_train_step(h, params_out->typed_data(), x.typed_data(), y.typed_data(), x.dimensions()[0],
                                  inv_global_b, nullptr, 1, loss->typed_data(), grad->typed_data(),
                                  m_out->typed_data(), v_out->typed_data(), step_out->typed_data(), &o, 0, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MafTrainStep, MafTrainStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<float>("inv_global_b")
                                  .Attr<float>("lr").Attr<float>("warmup").Attr<float>("decay_steps")
                                  .Attr<float>("clip").Attr<float>("weight_decay").Attr<int32_t>("kind")
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());
#endif  // xla/ffi/api/ffi.h
