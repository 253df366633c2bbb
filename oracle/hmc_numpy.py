"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the leapfrog integrator that mafb200_hmc_trajectory runs.

PARITY UNPINNED: jaxili delegates MCMC to numpyro (jaxili/posterior/mcmc_posterior.py:179-305), whose source is
not under /root/reference and which cannot be imported here; the reference's tests pin only output shapes
(jaxili/tests/test_posterior.py:147-166).  What is restated is the published algorithm numpyro's HMC runs:
velocity-Verlet leapfrog on the potential  -[log_prior(theta) + log_lik(x|theta) + log|d theta/d z|]  in the
unconstrained space of the prior's support (sigmoid bijection for interval supports, identity for the real
line), diagonal mass matrix, Metropolis correction on the energy error.

Only tests/ may import this module."""

import numpy as np


def constrain(z, kind, lo, hi):
    s = 1.0 / (1.0 + np.exp(-z))
    return np.where(kind == 1, lo + (hi - lo) * s, z)


def log_density_and_grad(z, kind, lo, hi, loglik_and_grad):
    """log density in z (up to a constant) and its gradient.  ``loglik_and_grad(theta) -> (ll[n], dll[n,C])``."""
    theta = constrain(z, kind, lo, hi)
    ll, dll = loglik_and_grad(theta)
    s = 1.0 / (1.0 + np.exp(-z))
    box_t = -np.abs(z) - 2.0 * np.log1p(np.exp(-np.abs(z)))
    box_g = dll * (hi - lo) * s * (1.0 - s) + (1.0 - 2.0 * s)
    d = (z - lo) / hi
    nor_t = -0.5 * d * d
    nor_g = dll - d / hi
    t = np.where(kind == 1, box_t, nor_t)
    g = np.where(kind == 1, box_g, nor_g)
    return ll + t.sum(axis=1), g, theta


def leapfrog(z, p, g, eps, inv_mass, n_steps, kind, lo, hi, loglik_and_grad):
    """n_steps of velocity Verlet; returns (z, p, g, logdens, theta)."""
    z, p, g = z.copy(), p.copy(), g.copy()
    ld, theta = None, None
    for _ in range(n_steps):
        p = p + 0.5 * eps * g
        z = z + eps * inv_mass * p
        ld, g, theta = log_density_and_grad(z, kind, lo, hi, loglik_and_grad)
        p = p + 0.5 * eps * g
    return z, p, g, ld, theta
