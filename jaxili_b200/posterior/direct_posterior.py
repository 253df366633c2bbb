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

    def sample_to_host(self, num_samples: int, key, x=None, out=None, chunk_rows: int = 1 << 18):
        """``sample`` delivered to host memory: the draw is cut into chunks whose device->host copies (into
        pinned memory, on their own stream) overlap the inversion kernel of the next chunk.  Identical values to
        ``sample(num_samples, key, x)`` (the Philox counter is the global row index).  Returns a pinned CPU
        tensor [num_samples, dim_theta]; ``out`` may supply it."""
        from ..model import as_device_f32, key_to_seed
        x = self._x(x)
        eng = self.model.engine()
        flat = self.model.flat_params(self.state.params)
        n_cond = eng.n_cond
        yo = as_device_f32(x, eng.device, n_cond) if n_cond else None
        if yo is not None and yo.shape[0] != 1:
            raise ValueError("sample_to_host conditions on a single observation x of shape (1, dim_x)")
        N, D = int(num_samples), eng.n_in
        if out is None:
            out = torch.empty((N, D), dtype=torch.float32).pin_memory()
        if tuple(out.shape) != (N, D) or out.dtype != torch.float32 or not out.is_pinned():
            raise ValueError("out must be a pinned float32 CPU tensor of shape (num_samples, dim_theta)")
        chunk = max(1, min(int(chunk_rows), N))
        bufs = [torch.empty((chunk, D), dtype=torch.float32, device=eng.device) for _ in range(2)]
        cur = torch.cuda.current_stream(eng.device)
        d2h = getattr(self, "_d2h_stream", None)
        if d2h is None:
            d2h = self._d2h_stream = torch.cuda.Stream(device=eng.device)
        done = [torch.cuda.Event(), torch.cuda.Event()]
        copied = [None, None]
        seed = key_to_seed(key)
        for k, r0 in enumerate(range(0, N, chunk)):
            n = min(chunk, N - r0)
            b = k & 1
            if copied[b] is not None:
                cur.wait_event(copied[b])                    # the buffer's previous chunk has left the device
            eng.sample_philox(flat, n, yo, seed=seed, row_offset=r0, rows_per_obs=max(n, 1), out=bufs[b][:n],
                              packed_current=k > 0)
            done[b].record(cur)
            d2h.wait_event(done[b])
            with torch.cuda.stream(d2h):
                out[r0:r0 + n].copy_(bufs[b][:n], non_blocking=True)
                copied[b] = torch.cuda.Event()
                copied[b].record(d2h)
        d2h.synchronize()
        return out

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
