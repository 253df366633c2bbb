"""Posterior objects (jaxili.posterior)."""

from .base_posterior import NeuralPosterior
from .direct_posterior import DirectPosterior
from .mcmc_posterior import MCMCPosterior
