"""Activation handles: stand-ins for the ``jax.nn`` callables jaxili passes around.

The reference stores the activation as a callable (``jax.nn.relu``, jaxili/inference/npe.py:42)
and serialises it by ``__name__`` (jaxili/train.py:208-211, inventory/func_dict.py:28-31).  The
fused kernels implement the activation themselves, so only the *name* matters here: any callable
whose ``__name__`` is a supported jax.nn name (including the real ``jax.nn.relu`` if JAX happens to
be installed) or the name itself is accepted.  The callables below also evaluate on torch tensors
so they can be used for ad-hoc checks.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

from ._lib import ACTIVATIONS


def relu(x):
    return torch.relu(x)


def silu(x):
    return F.silu(x)


swish = silu


def tanh(x):
    return torch.tanh(x)


def gelu(x):
    return F.gelu(x, approximate="tanh")


def elu(x):
    return F.elu(x)


def sigmoid(x):
    return torch.sigmoid(x)


def softplus(x):
    return F.softplus(x)


def leaky_relu(x):
    return F.leaky_relu(x, 0.01)


# name -> callable, the analogue of jaxili.inventory.func_dict.jax_nn_dict
jax_nn_dict = {f.__name__: f for f in (relu, silu, tanh, gelu, elu, sigmoid, softplus, leaky_relu)}
jax_nn_dict["swish"] = silu


def activation_name(act):
    """Resolve a callable / string to the kernel's activation name; anything else is an error."""
    name = act if isinstance(act, str) else getattr(act, "__name__", None)
    if name not in ACTIVATIONS:
        raise ValueError(
            f"activation {act!r} is not supported by the fused MAF kernels "
            f"(supported: {sorted(ACTIVATIONS)})")
    return "silu" if name == "swish" else name
