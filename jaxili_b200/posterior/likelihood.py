"""Likelihood evaluator for NLE-trained flows.

The reference's ``MCMCPosterior`` (jaxili/posterior/mcmc_posterior.py) wraps numpyro NUTS/HMC around
``log_likelihood(x, theta) = model.apply(params, x, theta, method="log_prob")`` (139-159).  numpyro is
not part of this hot path; what is: that log-likelihood call, evaluated here by the fused kernel for
whole batches of theta (many chains at once)."""

import numpy as np
import torch

from .base_posterior import NeuralPosterior


class LikelihoodEvaluator(NeuralPosterior):
    def __init__(self, model, state, verbose=False, x=None, prior_distr=None):
        super().__init__(model, state, verbose, x)
        self.prior_distr = prior_distr

    def log_likelihood(self, x, theta):
        """jaxili/posterior/mcmc_posterior.py:139-159: log p(x | theta), x broadcast over theta rows."""
        x = x.reshape(1, -1) if len(x.shape) == 1 else x
        theta = theta.reshape(1, -1) if len(theta.shape) == 1 else theta
        if x.shape[0] == 1 and theta.shape[0] > 1:
            x = x.expand(theta.shape[0], -1) if isinstance(x, torch.Tensor) else np.repeat(x, theta.shape[0], axis=0)
        return self.model.apply({"params": self.state.params}, x, theta, method="log_prob")

    def log_likelihood_and_grad(self, x, theta):
        """(log p(x | theta), d log p(x | theta) / d theta) for a batch of theta (many chains at once):
        the potential-energy gradient NUTS/HMC evaluates per leapfrog step
        (jaxili/posterior/mcmc_posterior.py:194-201), from one fused forward+reverse kernel launch."""
        x = x.reshape(1, -1) if len(x.shape) == 1 else x
        theta = theta.reshape(1, -1) if len(theta.shape) == 1 else theta
        if x.shape[0] == 1 and theta.shape[0] > 1:
            x = x.expand(theta.shape[0], -1) if isinstance(x, torch.Tensor) else np.repeat(x, theta.shape[0], axis=0)
        return self.model.apply({"params": self.state.params}, x, theta, method="log_prob_and_grad_cond")

    def unnormalized_log_prob(self, theta, x=None):
        x = self.x if x is None else x
        if x is None:
            raise ValueError(
                "Please set the default data `x` using `set_default_x()` or provide `x` as an argument.")
        ll = self.log_likelihood(x, theta)
        if self.prior_distr is not None and hasattr(self.prior_distr, "log_prob"):
            lp = self.prior_distr.log_prob(theta)
            ll = ll + torch.as_tensor(np.asarray(lp), device=ll.device, dtype=ll.dtype)
        return ll

    def sample(self, num_samples, key=None, x=None, mcmc_method=None, mcmc_kwargs=None):
        raise NotImplementedError(
            "MCMC sampling (numpyro NUTS/HMC in jaxili) is outside the B200 hot path; drive "
            "`unnormalized_log_prob` from an external sampler")
