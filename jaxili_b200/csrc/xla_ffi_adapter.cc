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
#endif  // xla/ffi/api/ffi.h
