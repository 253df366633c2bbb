"""Model shims: jaxili's ``jaxili.model`` surface for the MAF path, routed to the fused kernels.

Mirrors (names, arguments, ``method=`` strings, error behaviour):

* ``Identity``                 jaxili/model.py:25-43
* ``Standardizer``             jaxili/model.py:46-67
* ``NDENetwork``               jaxili/model.py:70-119
* ``ConditionalMAF``           jaxili/model.py:797-947
* ``NDE_w_Standardization``    jaxili/model.py:1059-1126
* ``ScalarAffine``             distrax.ScalarAffine as used in jaxili/inference/npe.py:391
* ``Sequential``               flax ``nn.Sequential`` as used in jaxili/inference/npe.py:401

flax modules are stateless descriptions used through ``init`` / ``apply`` / ``bind``; the same
three calls exist here.  Parameters are a flax-layout tree (jaxili_b200.layout) whose leaves are
views into one flat FP32 CUDA buffer; every numeric method runs in the native kernels -- there is
no PyTorch/CPU implementation of the flow to fall back to.

Arrays may be given as numpy arrays or torch tensors; results are torch CUDA tensors.
``key`` (a ``jax.random.PRNGKey`` in the reference) may be an int, a 2-element uint32 array or
anything ``np.asarray`` understands; it seeds Philox4x32-10 (JAX's threefry stream cannot be
reproduced, parity for samples is defined on supplied base noise: ``sample_from_noise``).
"""

from __future__ import annotations

from abc import abstractmethod

import numpy as np
import torch

from . import layout
from .engine import MafEngine
from .nn import activation_name


# --------------------------------------------------------------------------- helpers
def default_device():
    if not torch.cuda.is_available():
        from ._lib import MafB200Error
        raise MafB200Error("no CUDA device: jaxili_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_device_f32(a, device, cols=None):
    """numpy / torch -> contiguous FP32 tensor on ``device`` (2-D if ``cols`` is given)."""
    if isinstance(a, torch.Tensor):
        t = a.to(device=device, dtype=torch.float32)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(device)
    if cols is not None:
        t = t.reshape(-1, cols)
    return t.contiguous()


def key_to_seed(key):
    """Fold a PRNG key-like object into a 64-bit Philox seed."""
    if key is None:
        return 0
    if isinstance(key, (int, np.integer)):
        return int(key) & (2**64 - 1)
    if isinstance(key, torch.Tensor):
        key = key.detach().cpu().numpy()
    a = np.asarray(key).astype(np.uint64).reshape(-1)
    seed = 0
    for v in a:
        seed = ((seed << 32) ^ int(v)) & (2**64 - 1)
    return seed


def _np(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy().astype(np.float32)
    return np.asarray(a, dtype=np.float32)


class _Module:
    """The slice of flax.linen.Module behaviour jaxili relies on: init / apply / bind."""

    _bound = None

    def init(self, rngs, *args, method=None, **kwargs):
        raise NotImplementedError

    def apply(self, variables, *args, method=None, **kwargs):
        fn = getattr(self, method) if isinstance(method, str) else (method or self.__call__)
        if hasattr(fn, "__func__") and not isinstance(method, str) and method is not None:
            fn = getattr(self, fn.__name__)
        prev = self._bound
        self._bound = variables
        try:
            return fn(*args, **kwargs)
        finally:
            self._bound = prev

    def bind(self, variables):
        import copy
        m = copy.copy(self)
        m._bound = variables
        return m

    def _params(self):
        if self._bound is None:
            raise ValueError("module is not bound: call through .apply(variables, ...) or .bind(variables)")
        return self._bound["params"] if "params" in self._bound else self._bound


# --------------------------------------------------------------------------- small modules
class Identity(_Module):
    """Identity transformation (jaxili/model.py:25-43)."""

    def __call__(self, x):
        return x

    def init(self, rngs, *args, **kwargs):
        return {"params": {}}


class Standardizer(_Module):
    """z-score: (x - mean) / std (jaxili/model.py:46-67)."""

    def __init__(self, mean, std):
        self.mean, self.std = _np(mean), _np(std)

    def __call__(self, x):
        dev = x.device if isinstance(x, torch.Tensor) else default_device()
        x = as_device_f32(x, dev)
        return (x - as_device_f32(self.mean, dev)) / as_device_f32(self.std, dev)

    def init(self, rngs, *args, **kwargs):
        return {"params": {}}


class Sequential(_Module):
    """flax ``nn.Sequential`` stand-in (embedding_net = Sequential([standardizer, embedding]))."""

    def __init__(self, layers):
        self.layers = list(layers)

    def __call__(self, x):
        for l in self.layers:
            x = l(x)
        return x


class ScalarAffine:
    """distrax.ScalarAffine: forward y = scale*x + shift; inverse (y - shift) * (1/scale);
    inverse_log_det = -log|scale| per element."""

    def __init__(self, shift, scale):
        self.shift, self.scale = _np(shift), _np(scale)
        self._inv_scale = (np.float32(1.0) / self.scale).astype(np.float32)

    def forward(self, x):
        dev = x.device if isinstance(x, torch.Tensor) else default_device()
        return as_device_f32(self.scale, dev) * as_device_f32(x, dev) + as_device_f32(self.shift, dev)

    def inverse(self, y):
        dev = y.device if isinstance(y, torch.Tensor) else default_device()
        return (as_device_f32(y, dev) - as_device_f32(self.shift, dev)) * as_device_f32(self._inv_scale, dev)

    def inverse_and_log_det(self, y):
        x = self.inverse(y)
        ld = -torch.log(torch.abs(as_device_f32(self.scale, x.device)))
        return x, ld.expand_as(x)


def _standardizer_stats(embedding_net, n_cond):
    """(mean, std) of an embedding net made of Standardizer / Identity pieces; anything else is
    outside the fused path (the reference default is Identity, jaxili/inference/npe.py:346)."""
    mean, std = np.zeros(n_cond, np.float32), np.ones(n_cond, np.float32)
    pieces = embedding_net.layers if isinstance(embedding_net, Sequential) else [embedding_net]
    seen = False
    for p in pieces:
        if isinstance(p, Identity):
            continue
        if isinstance(p, Standardizer) and not seen:
            mean, std, seen = p.mean, p.std, True
            continue
        raise NotImplementedError(
            "only Standardizer/Identity embedding nets are fused into the B200 MAF path "
            f"(got {type(p).__name__}); compressor networks are out of scope")
    return mean, std


# --------------------------------------------------------------------------- NDE base
class NDENetwork(_Module):
    """Base class of a normalising flow (jaxili/model.py:70-119)."""

    @abstractmethod
    def log_prob(self, x, y=None, **kwargs):
        raise NotImplementedError("log_prob method not implemented in your child class of NDENetwork")

    @abstractmethod
    def sample(self, y, num_samples, key):
        raise NotImplementedError("sample method not implemented in tour child class of NDENetwork")


class ConditionalMAF(NDENetwork):
    """Conditional Masked Autoregressive Flow (jaxili/model.py:797-947), fused on B200.

    Same constructor fields as the flax dataclass: ``n_in, n_cond, n_layers, layers, activation,
    use_reverse, seed``.
    """

    def __init__(self, n_in, n_cond, n_layers, layers, activation=None, use_reverse=True, seed=None,
                 device=None):
        if activation is None:
            raise TypeError("ConditionalMAF missing required argument: 'activation'")
        self.n_in, self.n_cond, self.n_layers = int(n_in), int(n_cond), int(n_layers)
        self.layers = list(layers)
        self.activation = activation
        self.use_reverse = bool(use_reverse)
        self.seed = seed
        self._device = device
        self._act_name = activation_name(activation)
        self._engine = None
        self._std = (None, None, None, None)

    # ---- lazily created native handle (needs a GPU)
    @property
    def dims(self):
        return [self.n_in + self.n_cond, *self.layers, 2 * self.n_in]

    @property
    def n_params(self):
        return layout.n_params(self.dims, self.n_layers)

    def engine(self):
        if self._engine is None:
            dev = self._device or default_device()
            self._engine = MafEngine(self.n_in, self.n_cond, self.n_layers, self.layers, self._act_name,
                                     self.use_reverse, self.seed, *self._std, device=dev)
        return self._engine

    def with_standardization(self, shift, scale, mean, std):
        """Copy of this flow whose kernels fold the given standardisation constants."""
        import copy
        m = copy.copy(self)
        m._engine = None
        m._std = (shift, scale, mean, std)
        return m

    # ---- parameters
    def init(self, rngs, *args, method=None, **kwargs):
        """flax ``init``: truncated_normal(0.01) kernels (TN(-2,2)*0.01), zero biases
        (jaxili/model.py:532-537).  Only the shapes of ``args`` matter, as in the reference."""
        dev = self._device or default_device()
        gen = torch.Generator(device="cpu").manual_seed(key_to_seed(rngs) % (2**63))
        flat = torch.zeros(self.n_params, dtype=torch.float32)
        off = 0
        for _ in range(self.n_layers):
            for a, b in zip(self.dims[:-1], self.dims[1:]):
                w = torch.empty(a * b)
                torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
                flat[off:off + a * b] = w * 0.01
                off += a * b + b
        return {"params": layout.tree_from_flat(flat.to(dev), self.dims, self.n_layers)}

    def flat_params(self, params=None):
        eng = self.engine()
        tree = self._params() if params is None else (params["params"] if "params" in params else params)
        flat = layout.flat_from_tree(tree, self.dims, self.n_layers, device=eng.device)
        if flat.device != eng.device:
            flat = flat.to(eng.device)
        return flat.contiguous()

    # ---- numeric methods (all in the native kernels)
    def __call__(self, x, y=None):
        """(u, log_det_sum), jaxili/model.py:848-874."""
        eng = self.engine()
        return eng.forward(self.flat_params(), as_device_f32(x, eng.device, self.n_in),
                           self._cond(y, None, eng))

    def backward(self, u, y=None):
        """(x, log_det_sum) of the inverse flow, jaxili/model.py:876-901."""
        eng = self.engine()
        u = as_device_f32(u, eng.device, self.n_in)
        return eng.inverse(self.flat_params(), u, self._cond(y, u.shape[0], eng))

    def log_prob(self, x, y=None):
        """jaxili/model.py:903-921."""
        eng = self.engine()
        x = as_device_f32(x, eng.device, self.n_in)
        return eng.log_prob(self.flat_params(), x, self._cond(y, x.shape[0], eng))

    def sample(self, y=None, num_samples=1, key=None):
        """jaxili/model.py:923-947; ``y`` of shape (n_cond,) or (1, n_cond) is broadcast."""
        eng = self.engine()
        yo = as_device_f32(y, eng.device, self.n_cond) if self.n_cond else None
        return eng.sample_philox(self.flat_params(), int(num_samples), yo, seed=key_to_seed(key),
                                 rows_per_obs=int(num_samples) if yo is None or yo.shape[0] == 1 else None)

    def sample_from_noise(self, u, y=None):
        """Inverse flow of supplied base noise (parity entry point; not in the reference API)."""
        eng = self.engine()
        u = as_device_f32(u, eng.device, self.n_in)
        yo = as_device_f32(y, eng.device, self.n_cond) if self.n_cond else None
        return eng.sample_from_noise(self.flat_params(), u, yo,
                                     rows_per_obs=u.shape[0] if yo is None or yo.shape[0] == 1 else None)

    def _cond(self, y, rows, eng):
        if not self.n_cond:
            return None
        y = as_device_f32(y, eng.device, self.n_cond)
        if rows is not None and y.shape[0] == 1 and rows > 1:
            y = y.expand(rows, -1).contiguous()
        return y


class NDE_w_Standardization(NDENetwork):
    """NDE with standardisation of both variables (jaxili/model.py:1059-1126).

    ``nde`` must be a ConditionalMAF here; ``embedding_net`` a Standardizer / Identity (possibly in
    a Sequential); ``transformation`` a ScalarAffine.  The constants are folded into the kernels.
    """

    def __init__(self, nde, embedding_net, transformation):
        if not isinstance(nde, ConditionalMAF):
            raise NotImplementedError("the B200 path implements NDE_w_Standardization for ConditionalMAF only")
        self.nde = nde
        self.embedding_net = embedding_net
        self.transformation = transformation
        mean, std = _standardizer_stats(embedding_net, nde.n_cond)
        self._fused = nde.with_standardization(transformation.shift, transformation.scale, mean, std)

    @property
    def dims(self):
        return self.nde.dims

    @property
    def n_layers(self):
        return self.nde.n_layers

    def engine(self):
        return self._fused.engine()

    def init(self, rngs, *args, method=None, **kwargs):
        inner = self.nde.init(rngs)["params"]
        return {"params": layout.tree_from_flat(inner.flat, self.nde.dims, self.nde.n_layers, wrap_nde=True)}

    def flat_params(self, params=None):
        tree = self._params() if params is None else (params["params"] if "params" in params else params)
        eng = self.engine()
        flat = layout.flat_from_tree(tree, self.nde.dims, self.nde.n_layers, device=eng.device)
        return flat.to(eng.device).contiguous()

    def __call__(self, x, y, model="NPE"):
        assert model in ["NPE", "NLE"], "Model should be either 'NPE' or 'NLE'."
        if model == "NLE":
            x, y = y, x
        eng = self.engine()
        x = as_device_f32(x, eng.device, self.nde.n_in)
        y = self._fused._cond(y, x.shape[0], eng)
        return eng.log_prob(self.flat_params(), x, y)

    def standardize(self, x):
        return self.transformation.inverse(x)

    def unstandardize(self, x):
        return self.transformation.forward(x)

    def embedding(self, x):
        return self.embedding_net(as_device_f32(x, default_device()))

    def log_prob(self, x, y=None, model="NPE"):
        return self.__call__(x, y, model)

    def log_prob_and_grad_cond(self, x, y):
        """(log_prob, d log_prob / d y) -- the value and gradient an HMC/NUTS sampler over an NLE
        likelihood needs (jax.grad of jaxili/posterior/mcmc_posterior.py:139-159 w.r.t. theta)."""
        eng = self.engine()
        x = as_device_f32(x, eng.device, self.nde.n_in)
        y = self._fused._cond(y, x.shape[0], eng)
        return eng.log_prob_grad_y(self.flat_params(), x, y)

    def sample(self, y, num_samples, key, model="NPE"):
        assert model in ["NPE", "NLE"], "Model should be either 'NPE' or 'NLE'."
        eng = self.engine()
        yo = as_device_f32(y, eng.device, self.nde.n_cond) if self.nde.n_cond else None
        return eng.sample_philox(self.flat_params(), int(num_samples), yo, seed=key_to_seed(key),
                                 rows_per_obs=int(num_samples) if yo is None or yo.shape[0] == 1 else None)

    def sample_from_noise(self, u, y):
        eng = self.engine()
        u = as_device_f32(u, eng.device, self.nde.n_in)
        yo = as_device_f32(y, eng.device, self.nde.n_cond) if self.nde.n_cond else None
        return eng.sample_from_noise(self.flat_params(), u, yo,
                                     rows_per_obs=u.shape[0] if yo is None or yo.shape[0] == 1 else None)
