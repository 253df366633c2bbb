// Vectorised-chains Hamiltonian Monte Carlo around the fused flow kernel.
//
// Replaces the inner loop of jaxili's MCMCPosterior (jaxili/posterior/mcmc_posterior.py:179-305:
// numpyro HMC/NUTS over  log_prior(theta) + log_likelihood(x, theta),  139-201): every chain is one
// row, one leapfrog step of ALL chains is
//     hmc_drift_kernel  ->  maf_flow_kernel<BWD> in d/dy mode  ->  hmc_kick_kernel
// and a trajectory is 3 x n_steps launches enqueued by one C call.  The sampler works in the
// unconstrained space z of the prior's support (numpyro does the same: biject_to(support)):
//     kind 0  real line   theta = z                       prior Normal(lo = loc, hi = scale)
//     kind 1  interval    theta = lo + (hi - lo) sigmoid(z) prior Uniform(lo, hi)
// log density in z (up to a constant) = log_lik(theta) + log_prior(theta) + log |d theta / d z|.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mafb {

struct HmcArgs {
  const float* eps;        // [1] device scalar step size (adapted on the device between trajectories)
  const float* inv_mass;   // [C] diagonal inverse mass matrix
  const int* kind;         // [C]
  const float* lo;         // [C]
  const float* hi;         // [C]
  float* z;                // [n][C] position (unconstrained)
  float* p;                // [n][C] momentum
  float* g;                // [n][C] gradient of the log density at z
  float* theta;            // [n][C] constrained position, the flow's conditioning input
  const float* dll;        // [n][C] d log_lik / d theta from the flow kernel
  const float* ll;         // [n]    log_lik
  float* logdens;          // [n]    log density at z
  long long n;
  int C;
  float kick;              // momentum update p += kick * eps * g  (0.5 inside a trajectory, 0 to evaluate)
  float drift;             // position update z += drift * eps * inv_mass * p  (1 inside a trajectory)
};

// half kick with the old gradient, full drift, map to the constrained space
__global__ void hmc_drift_kernel(const HmcArgs A) {
  const long long total = A.n * A.C;
  const float eps = *A.eps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % A.C);
    float p = A.p[i], z = A.z[i];
    if (A.kick != 0.f) p = fmaf(A.kick * eps, A.g[i], p);
    if (A.drift != 0.f) z = fmaf(A.drift * eps * A.inv_mass[c], p, z);
    A.p[i] = p;
    A.z[i] = z;
    float th = z;
    if (A.kind[c] == 1) {
      // exact float sigmoid so that theta stays strictly inside (lo, hi) for |z| < 16
      const float s = 1.f / (1.f + expf(-z));
      th = fmaf(A.hi[c] - A.lo[c], s, A.lo[c]);
    }
    A.theta[i] = th;
  }
}

// chain rule to z, prior and Jacobian terms, second half kick, log density.  One warp per chain.
__global__ void hmc_kick_kernel(const HmcArgs A) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float eps = *A.eps;
  for (long long r = warp; r < A.n; r += n_warps) {
    float lp = 0.f;
    for (int c = lane; c < A.C; c += 32) {
      const long long i = r * A.C + c;
      const float z = A.z[i], dl = A.dll[i];
      float g, t;
      if (A.kind[c] == 1) {
        const float s = 1.f / (1.f + expf(-z));
        // log s + log(1 - s) = -|z| - 2 log(1 + exp(-|z|)), stable for large |z|
        t = -fabsf(z) - 2.f * log1pf(expf(-fabsf(z)));
        g = fmaf(dl, (A.hi[c] - A.lo[c]) * s * (1.f - s), 1.f - 2.f * s);
      } else {
        const float sc = A.hi[c], d = (z - A.lo[c]) / sc;
        t = -0.5f * d * d;
        g = dl - d / sc;
      }
      lp += t;
      A.g[i] = g;
      if (A.kick != 0.f) A.p[i] = fmaf(A.kick * eps, g, A.p[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
    if (lane == 0) A.logdens[r] = A.ll[r] + lp;
  }
}

}  // namespace mafb
