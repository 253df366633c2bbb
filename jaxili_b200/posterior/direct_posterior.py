"""DirectPosterior (jaxili/posterior/direct_posterior.py:16-107) on the fused kernels."""

from typing import Optional

import numpy as np
import torch

from .base_posterior import NeuralPosterior


class DirectPosterior(NeuralPosterior):
    r"""Posterior $p(\theta|x)$ of an NPE-trained flow: ``sample`` runs the autoregressive-inversion
    kernel, ``unnormalized_log_prob`` the fused forward kernel."""

    def __init__(self, model, state, verbose: bool = False, x=None):
        super().__init__(model, state, verbose, x)

    def _x(self, x):
        if x is None:
            x = self.x
            if x is None:
                raise ValueError(
                    "Please set the default data `x` using `set_default_x()` or provide `x` as an argument.")
        return x

    def sample(self, num_samples: int, key, x=None):
        """jaxili/posterior/direct_posterior.py:45-73."""
        x = self._x(x)
        return self.model.apply({"params": self.state.params}, x, num_samples, key, method="sample")

    def sample_from_noise(self, u, x=None):
        """Same inversion with the base noise supplied (parity entry point)."""
        x = self._x(x)
        return self.model.apply({"params": self.state.params}, u, x, method="sample_from_noise")

    def unnormalized_log_prob(self, theta, x=None):
        """jaxili/posterior/direct_posterior.py:75-107: one ``x`` is broadcast to theta's batch."""
        x = self._x(x)
        xs = x.shape if hasattr(x, "shape") else np.asarray(x).shape
        if len(xs) == 1:
            x = x.reshape(1, -1)
        if (x.shape[0] == 1) and (theta.shape[0] > 1):
            x = x.expand(theta.shape[0], -1) if isinstance(x, torch.Tensor) else np.repeat(x, theta.shape[0], axis=0)
        elif x.shape[0] != theta.shape[0]:
            raise ValueError("The batch size of `x` must be the same as the batch size of parameters `theta`.")
        return self.model.apply({"params": self.state.params}, theta, x, method="log_prob")
