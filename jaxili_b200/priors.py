"""Prior distributions for ``MCMCPosterior`` -- the two numpyro distributions jaxili's NLE examples and
tests pass as ``prior_distr`` (``dist.Uniform(low, high)``: jaxili/tests/test_posterior.py:123;
``dist.Normal``), with the slice of the numpyro interface MCMCPosterior touches
(jaxili/posterior/mcmc_posterior.py:120-137, 429-433): ``log_prob(theta)`` per element and
``sample(key, sample_shape)``.  Samples are torch CUDA tensors drawn with torch's Philox generator."""

import math

import numpy as np
import torch


def _gen(key, device):
    from .model import key_to_seed
    g = torch.Generator(device=device)
    g.manual_seed(key_to_seed(key) & (2**63 - 1))
    return g


class Distribution:
    kind = -1

    def _dev(self, device=None):
        from .model import default_device
        return device if device is not None else default_device()

    def bounds(self):
        """(prior_kind, lo, hi) float32 arrays of the C-ABI's MafHmcState convention."""
        raise NotImplementedError


class Uniform(Distribution):
    kind = 1

    def __init__(self, low, high):
        self.low = np.atleast_1d(np.asarray(low, np.float32))
        self.high = np.atleast_1d(np.asarray(high, np.float32))
        if self.low.shape != self.high.shape or not np.all(self.high > self.low):
            raise ValueError("Uniform needs low < high of equal shapes")

    def log_prob(self, theta):
        t = theta if isinstance(theta, torch.Tensor) else torch.as_tensor(np.asarray(theta, np.float32))
        lo = torch.as_tensor(self.low, device=t.device)
        hi = torch.as_tensor(self.high, device=t.device)
        inside = (t >= lo) & (t < hi)
        return torch.where(inside, -torch.log(hi - lo).expand_as(t), torch.full_like(t, -math.inf))

    def sample(self, key=None, sample_shape=(), device=None):
        dev = self._dev(device)
        u = torch.rand(tuple(sample_shape) + self.low.shape, generator=_gen(key, dev), device=dev)
        lo, hi = torch.as_tensor(self.low, device=dev), torch.as_tensor(self.high, device=dev)
        return lo + (hi - lo) * u

    def bounds(self):
        return np.full(self.low.shape, 1, np.int32), self.low, self.high


class Normal(Distribution):
    kind = 0

    def __init__(self, loc=0.0, scale=1.0):
        loc, scale = np.broadcast_arrays(np.atleast_1d(np.asarray(loc, np.float32)),
                                         np.atleast_1d(np.asarray(scale, np.float32)))
        self.loc, self.scale = np.ascontiguousarray(loc), np.ascontiguousarray(scale)
        if not np.all(self.scale > 0):
            raise ValueError("Normal needs scale > 0")

    def log_prob(self, theta):
        t = theta if isinstance(theta, torch.Tensor) else torch.as_tensor(np.asarray(theta, np.float32))
        loc = torch.as_tensor(self.loc, device=t.device)
        sc = torch.as_tensor(self.scale, device=t.device)
        return -0.5 * ((t - loc) / sc) ** 2 - torch.log(sc) - 0.5 * math.log(2 * math.pi)

    def sample(self, key=None, sample_shape=(), device=None):
        dev = self._dev(device)
        e = torch.randn(tuple(sample_shape) + self.loc.shape, generator=_gen(key, dev), device=dev)
        return torch.as_tensor(self.loc, device=dev) + torch.as_tensor(self.scale, device=dev) * e

    def bounds(self):
        return np.zeros(self.loc.shape, np.int32), self.loc, self.scale
