"""NeuralPosterior base class (jaxili/posterior/base_posterior.py:16-69)."""

from abc import abstractmethod
from typing import Any, Dict, Optional


class NeuralPosterior:
    r"""Posterior $p(\theta|x)$ with ``log_prob()`` and ``sample()`` methods."""

    def __init__(self, model, state, verbose: bool = False, x=None):
        self.model = model
        self.state = state
        self.verbose = verbose
        self.x = x

    @abstractmethod
    def sample(self, num_samples: int, key, x, mcmc_method: Optional[str] = None,
               mcmc_kwargs: Optional[Dict[str, Any]] = None):
        pass

    @abstractmethod
    def unnormalized_log_prob(self, theta):
        pass

    def set_default_x(self, x):
        """Set the default data for the posterior."""
        self.x = x
