"""Independent torch-CPU restatement of the jaxili MAF hot path (autograd).

TEST INFRASTRUCTURE ONLY (see oracle/maf_numpy.py header: same rules, same
"PARITY UNPINNED" caveat).  Two jobs:

1. cross-check the hand-derived gradients of oracle/maf_numpy.py with reverse-mode
   autodiff, the way the reference obtains them (``jax.value_and_grad``,
   jaxili/train.py:330-332);
2. serve as the multi-threaded CPU baseline that ``bench.py`` times beside the GPU
   numbers (``cpu_baseline.kind = "port"``): the reference's own CPU path is XLA's
   multi-threaded FP32 dot + elementwise ops; torch's CPU matmul is the closest
   thing that runs in this image.

Written against the reference lines directly (not against maf_numpy.py) so that the
two restatements are independent apart from the mask builder.
"""

from __future__ import annotations

import math

import numpy as np
import torch

from . import maf_numpy as onp

LOG_2PI = math.log(2.0 * math.pi)


def _act(name):
    F = torch.nn.functional
    return {
        "relu": torch.relu,
        "silu": F.silu,
        "swish": F.silu,
        "tanh": torch.tanh,
        "sigmoid": torch.sigmoid,
        "elu": F.elu,
        "leaky_relu": lambda a: F.leaky_relu(a, 0.01),
        "softplus": F.softplus,
        "gelu": lambda a: F.gelu(a, approximate="tanh"),
    }[name]


class TorchMaf:
    """ConditionalMAF + NDE_w_Standardization on torch CPU tensors."""

    def __init__(self, spec: onp.MafSpec, std: onp.Standardization, dtype=torch.float32):
        self.spec, self.dtype = spec, dtype
        self.act = _act(spec.activation)
        self.masks = [[torch.tensor(m, dtype=dtype) for m in ms] for ms in spec.masks]
        t = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
        self.shift, self.scale, self.mean, self.std = t(std.shift), t(std.scale), t(std.mean), t(std.std)

    def unflatten(self, flat):
        s, out, off = self.spec, [], 0
        for _ in range(s.n_layers):
            lay = []
            for i in range(s.n_lin):
                a, b = s.dims[i], s.dims[i + 1]
                lay.append((flat[off:off + a * b].view(a, b), flat[off + a * b:off + a * b + b]))
                off += a * b + b
            out.append(lay)
        return out

    def made(self, lay, masks, x, y):
        # jaxili/model.py:681-684 concat, 544 masked dot, 592-596 act between linears
        h = torch.cat([x, y], dim=1) if self.spec.n_cond else x
        n = self.spec.n_lin
        for i in range(n):
            W, b = lay[i]
            h = h @ (masks[i] * W) + b
            if i < n - 1:
                h = self.act(h)
        return h

    def maf_log_prob(self, flat, x, y):
        # jaxili/model.py:869-874, 738-743, 919-921
        s = self.spec
        P = self.unflatten(flat)
        D = s.n_in
        ld = torch.zeros(x.shape[0], dtype=self.dtype)
        for k in range(s.n_layers):
            out = self.made(P[k], self.masks[k], x, y)
            mu, logp = out[:, :D], out[:, D:]
            u = (x - mu) * torch.exp(0.5 * logp)
            x = torch.flip(u, dims=[1]) if s.use_reverse else u
            ld = ld + 0.5 * logp.sum(dim=1)
        return -0.5 * (x * x).sum(dim=1) - 0.5 * D * LOG_2PI + ld

    def log_prob(self, flat, x, y):
        # jaxili/model.py:1092-1098
        xs = (x - self.shift) * (1.0 / self.scale)
        c = (-torch.log(torch.abs(self.scale))).sum()
        ys = (y - self.mean) / self.std
        return self.maf_log_prob(flat, xs, ys) + c

    def loss(self, flat, x, y):
        # jaxili/loss.py:57-58 / 83
        return -self.log_prob(flat, x, y).mean()

    def loss_and_grad(self, flat_np, x_np, y_np, want_dy=False):
        flat = torch.tensor(flat_np, dtype=self.dtype, requires_grad=True)
        x = torch.tensor(x_np, dtype=self.dtype)
        y = torch.tensor(y_np, dtype=self.dtype, requires_grad=want_dy)
        loss = self.loss(flat, x, y)
        loss.backward()
        if want_dy:
            return loss.item(), flat.grad.numpy(), y.grad.numpy()
        return loss.item(), flat.grad.numpy()

    @torch.no_grad()
    def sample_from_noise(self, flat_np, u_np, y_obs):
        # jaxili/model.py:1116-1126, 941-947, 896-901, 765-773
        s = self.spec
        flat = torch.tensor(flat_np, dtype=self.dtype)
        u = torch.tensor(u_np, dtype=self.dtype)
        y = torch.tensor(np.asarray(y_obs), dtype=self.dtype).reshape(-1, s.n_cond)
        y = (y - self.mean) / self.std
        y = y * torch.ones((u.shape[0], 1), dtype=self.dtype)
        P = self.unflatten(flat)
        D = s.n_in
        for k in reversed(range(s.n_layers)):
            if s.use_reverse:
                u = torch.flip(u, dims=[1])
            x = torch.zeros_like(u)
            for d in range(D):
                out = self.made(P[k], self.masks[k], x, y)
                mu, logp = out[:, :D], out[:, D:]
                m = torch.clamp(-0.5 * logp, max=10.0)
                x[:, d] = mu[:, d] + torch.exp(m[:, d]) * u[:, d]
            u = x
        return (self.scale * u + self.shift).numpy()


class TorchTrainer:
    """train_step of jaxili/train.py:330-335 with the optimiser chain of 246-300,
    on torch CPU: the timed CPU baseline.  Adam is hand-applied (optax formula) so
    the LR schedule and global-norm clip follow the reference exactly."""

    def __init__(self, model: TorchMaf, flat_np, lr=5e-4, warmup=0.1, decay_steps=10**9,
                 clip=5.0, b1=0.9, b2=0.999, eps=1e-8):
        self.model = model
        self.flat = torch.tensor(flat_np, dtype=model.dtype, requires_grad=True)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.count = 0
        self.h = dict(lr=lr, warmup=warmup, decay_steps=decay_steps, clip=clip, b1=b1, b2=b2, eps=eps)

    def train_step(self, x, y):
        h = self.h
        if self.flat.grad is not None:
            self.flat.grad = None
        loss = self.model.loss(self.flat, x, y)
        loss.backward()
        with torch.no_grad():
            g = self.flat.grad
            gn = torch.sqrt((g * g).sum())
            if not bool(gn < h["clip"]):
                g = (g / gn) * h["clip"]
            t = self.count
            self.m.mul_(h["b1"]).add_(g, alpha=1 - h["b1"])
            self.v.mul_(h["b2"]).addcmul_(g, g, value=1 - h["b2"])
            mhat = self.m / (1 - h["b1"] ** (t + 1))
            vhat = self.v / (1 - h["b2"] ** (t + 1))
            lr_t = onp.warmup_cosine_lr(t, h["lr"], h["warmup"], h["decay_steps"])
            self.flat.sub_(lr_t * mhat / (torch.sqrt(vhat) + h["eps"]))
            self.count = t + 1
        return loss.item()
