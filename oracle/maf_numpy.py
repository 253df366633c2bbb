"""CPU oracle for the jaxili ConditionalMAF hot path (NumPy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``jaxili_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the reported CPU baseline, never as the product path.

PARITY UNPINNED.  The reference (sachaguer/jaxili) is pure Python on
jax/flax/optax/distrax, none of which is installed in this image (no wheels, no
network), so jaxili itself cannot be imported here to generate vectors, and its
own tests hold no numeric golden values for this path (only the round-trip
invariant, jaxili/tests/test_nn.py:35-41).  This file restates the reference's
algorithm line by line; what *is* pinned without JAX is checked in
tests/test_oracle.py: the legacy-NumPy-RNG layer seeds and mask matrices
(jaxili/model.py:583,603-663,831-843), the forward/inverse round trip at 1e-5,
analytic gradients vs finite differences and vs an independent torch-autograd
restatement (oracle/maf_torch.py), and the optimiser chain vs torch.optim.Adam.

Every function cites the reference lines it follows (paths relative to
/root/reference/).  ``dtype`` selects float64 (ground truth) or float32 (the
rounding envelope the FP32 kernels are expected to sit in).

Flat parameter order used everywhere in this repo (oracle, C-ABI, host shims):
for MAF layer k = 0..L-1, for masked-linear i = 0..n_lin-1: kernel_i of shape
(in_i, out_i) row-major (flax ``Dense`` convention, jaxili/model.py:532-537),
then bias_i of shape (out_i,).
"""

from __future__ import annotations

import math

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------
# activations (jax.nn.* semantics; jaxili/inference/npe.py:42 default relu,
# jaxili/tests/test_nn.py:24 uses silu, inventory/func_dict.py:28-31 any jax.nn)
# --------------------------------------------------------------------------
def _sigmoid(a):
    return 1.0 / (1.0 + np.exp(-a))


def act_forward(name, a):
    """Return (h, dh/da) for activation ``name``; relu'(0) = 0 as in jax.nn.relu."""
    if name == "relu":
        return np.maximum(a, 0), (a > 0).astype(a.dtype)
    if name == "silu" or name == "swish":
        s = _sigmoid(a)
        return a * s, s * (1 + a * (1 - s))
    if name == "tanh":
        t = np.tanh(a)
        return t, 1 - t * t
    if name == "sigmoid":
        s = _sigmoid(a)
        return s, s * (1 - s)
    if name == "elu":
        e = np.exp(np.minimum(a, 0)) - 1
        return np.where(a > 0, a, e), np.where(a > 0, 1.0, e + 1).astype(a.dtype)
    if name == "leaky_relu":
        return np.where(a >= 0, a, 0.01 * a), np.where(a >= 0, 1.0, 0.01).astype(a.dtype)
    if name == "softplus":
        return np.logaddexp(a, 0), _sigmoid(a)
    if name == "gelu":  # jax.nn.gelu default approximate=True (tanh form)
        c = math.sqrt(2.0 / math.pi)
        inner = c * (a + 0.044715 * a**3)
        t = np.tanh(inner)
        dinner = c * (1 + 3 * 0.044715 * a * a)
        return 0.5 * a * (1 + t), 0.5 * (1 + t) + 0.5 * a * (1 - t * t) * dinner
    raise ValueError(f"unknown activation {name!r}")


# --------------------------------------------------------------------------
# A.1  masks and layer seeds
# --------------------------------------------------------------------------
def layer_seeds(seed, n_layers):
    """jaxili/model.py:831-843: ``np.random.seed(seed)`` then one
    ``np.random.randint(0, 1000)`` per layer, all drawn before any MADE setup."""
    np.random.seed(seed)
    return [int(np.random.randint(0, 1000)) for _ in range(n_layers)]


def made_degrees(n_in, hidden_dims, seed):
    """jaxili/model.py:583,610-619: degrees m^0..m^{n+1} (natural input order)."""
    np.random.seed(seed)
    D = n_in
    deg = [np.arange(D)]
    for h in hidden_dims:
        low = deg[-1].min()
        deg.append(np.random.randint(low, D - 1, size=h))
    deg.append(deg[0])
    return deg


def made_masks(n_in, n_cond, hidden_dims, seed):
    """jaxili/model.py:603-663: list of (in,out) 0/1 matrices, one per masked linear.

    First matrix has the C all-ones conditioning rows appended below the D data
    rows (622-629); the last one is strict and duplicated for (mu, logp) (644-657).
    """
    deg = made_degrees(n_in, hidden_dims, seed)
    C = n_cond
    mats = []
    m, m_next = deg[0], deg[1]
    M = np.ones((len(m), len(m_next)))
    for j in range(len(m_next)):
        M[:, j] = (m <= m_next[j]).astype(int)
    M = np.concatenate([M, np.ones((C, len(m_next)))], axis=0)
    mats.append(M)
    for i in range(1, len(deg) - 2):
        m, m_next = deg[i], deg[i + 1]
        M = np.zeros((len(m), len(m_next)))
        for j in range(len(m_next)):
            M[:, j] = (m <= m_next[j]).astype(int)
        mats.append(M)
    m, m_next = deg[len(deg) - 2], deg[len(deg) - 1]
    M = np.zeros((len(m), len(m_next)))
    for j in range(len(m)):
        M[j, :] = (m[j] < m_next).astype(int)
    mats.append(np.concatenate([M, M], axis=1))
    return mats


class MafSpec:
    """Static description of a ConditionalMAF (jaxili/model.py:821-846)."""

    def __init__(self, n_in, n_cond, n_layers, layers, activation="relu",
                 use_reverse=True, seed=42):
        self.n_in, self.n_cond, self.n_layers = int(n_in), int(n_cond), int(n_layers)
        self.layers = [int(h) for h in layers]
        self.activation = activation
        self.use_reverse = bool(use_reverse)
        self.seed = seed
        self.dims = [self.n_in + self.n_cond, *self.layers, 2 * self.n_in]
        self.seeds = layer_seeds(seed, self.n_layers)
        self.masks = [made_masks(self.n_in, self.n_cond, self.layers, s) for s in self.seeds]

    @property
    def n_lin(self):
        return len(self.dims) - 1

    @property
    def params_per_layer(self):
        return sum(self.dims[i] * self.dims[i + 1] + self.dims[i + 1] for i in range(self.n_lin))

    @property
    def n_params(self):
        return self.n_layers * self.params_per_layer

    def unflatten(self, flat):
        """flat -> [[(W_i, b_i) for i] for k] (views)."""
        out, off = [], 0
        for _ in range(self.n_layers):
            lay = []
            for i in range(self.n_lin):
                a, b = self.dims[i], self.dims[i + 1]
                W = flat[off:off + a * b].reshape(a, b)
                off += a * b
                bias = flat[off:off + b]
                off += b
                lay.append((W, bias))
            out.append(lay)
        assert off == flat.size
        return out

    def flat_mask(self):
        """0/1 vector in flat parameter order; biases are 1."""
        parts = []
        for k in range(self.n_layers):
            for i in range(self.n_lin):
                parts.append(self.masks[k][i].reshape(-1))
                parts.append(np.ones(self.dims[i + 1]))
        return np.concatenate(parts)

    def init_params(self, rng, scale=0.01, dtype=np.float32):
        """flax ``truncated_normal(0.01)`` kernels (TN(-2,2)*stddev), zero biases
        (jaxili/model.py:532-537).  The RNG stream is NOT jax's; only the
        distribution is restated (parity is defined on given params)."""
        flat = np.zeros(self.n_params, dtype=dtype)
        off = 0
        for _ in range(self.n_layers):
            for i in range(self.n_lin):
                a, b = self.dims[i], self.dims[i + 1]
                z = rng.standard_normal(a * b)
                bad = np.abs(z) > 2
                while bad.any():
                    z[bad] = rng.standard_normal(int(bad.sum()))
                    bad = np.abs(z) > 2
                flat[off:off + a * b] = (z * scale).astype(dtype)
                off += a * b + b
        return flat


# --------------------------------------------------------------------------
# A.2  forward / log_prob
# --------------------------------------------------------------------------
def made_forward(spec, lay, masks, xin, y, keep=False):
    """jaxili/model.py:665-686 + 517-544: concat([x,y]) -> (mask*W) dots with act."""
    h = np.concatenate([xin, y], axis=1) if spec.n_cond else xin
    hs, ds = [h], []
    n = spec.n_lin
    for i in range(n):
        W, b = lay[i]
        a = h @ (masks[i].astype(W.dtype) * W) + b
        if i < n - 1:
            h, d = act_forward(spec.activation, a)
            hs.append(h)
            ds.append(d)
        else:
            h = a
    return (h, hs, ds) if keep else h


def maf_forward(spec, flat, x, y, keep=False):
    """jaxili/model.py:848-874 with MAFLayer.forward 718-743.  Returns u, log_det
    (and the per-layer tape when ``keep``)."""
    dt = flat.dtype
    x = x.astype(dt)
    y = y.astype(dt)
    P = spec.unflatten(flat)
    D = spec.n_in
    ld = np.zeros(x.shape[0], dtype=dt)
    tape = []
    for k in range(spec.n_layers):
        if keep:
            out, hs, ds = made_forward(spec, P[k], spec.masks[k], x, y, keep=True)
        else:
            out = made_forward(spec, P[k], spec.masks[k], x, y)
        mu, logp = out[:, :D], out[:, D:]
        s = np.exp(dt.type(0.5) * logp)
        v = (x - mu) * s
        if keep:
            tape.append((x, hs, ds, s, v))
        x = v[:, ::-1] if spec.use_reverse else v
        ld = ld + dt.type(0.5) * logp.sum(axis=1)
    return (x, ld, tape) if keep else (x, ld)


def maf_log_prob(spec, flat, x, y):
    """jaxili/model.py:903-921: N(0,I) logpdf of u plus the summed log-dets."""
    u, ld = maf_forward(spec, flat, x, y)
    dt = flat.dtype
    return dt.type(-0.5) * (u * u).sum(axis=1) - dt.type(0.5 * spec.n_in * LOG_2PI) + ld


class Standardization:
    """Constants of NDE_w_Standardization (jaxili/model.py:1073-1098): ``shift``,
    ``scale`` of the distrax.ScalarAffine acting on the density variable and
    ``mean``, ``std`` of the Standardizer acting on the conditioning variable
    (jaxili/inference/npe.py:383-399, nle.py:323-341)."""

    def __init__(self, shift, scale, mean, std):
        self.shift, self.scale, self.mean, self.std = [np.asarray(a) for a in (shift, scale, mean, std)]

    @staticmethod
    def identity(n_in, n_cond):
        return Standardization(np.zeros(n_in), np.ones(n_in), np.zeros(n_cond), np.ones(n_cond))


def nde_log_prob(spec, flat, std, x, y):
    """jaxili/model.py:1092-1098: x'=(x-shift)*(1/scale), c=sum(-log|scale|),
    y'=(y-mean)/std, result = MAF.log_prob(x',y') + c."""
    dt = flat.dtype
    xs = (x.astype(dt) - std.shift.astype(dt)) * (dt.type(1.0) / std.scale.astype(dt))
    c = (-np.log(np.abs(std.scale.astype(dt)))).sum()
    ys = (y.astype(dt) - std.mean.astype(dt)) / std.std.astype(dt)
    return maf_log_prob(spec, flat, xs, ys) + c


# --------------------------------------------------------------------------
# A.3  inverse / sample
# --------------------------------------------------------------------------
def maf_inverse(spec, flat, u, y):
    """jaxili/model.py:876-901 with MAFLayer.backward 745-773: D full MADE passes
    per layer, clamp min(-0.5*logp, 10), log-det of the last pass only."""
    dt = flat.dtype
    u = u.astype(dt)
    y = y.astype(dt)
    P = spec.unflatten(flat)
    D = spec.n_in
    ld_sum = np.zeros(u.shape[0], dtype=dt)
    for k in reversed(range(spec.n_layers)):
        if spec.use_reverse:
            u = u[:, ::-1]
        x = np.zeros_like(u)
        for d in range(D):
            out = made_forward(spec, P[k], spec.masks[k], x, y)
            mu, logp = out[:, :D], out[:, D:]
            m = np.minimum(dt.type(-0.5) * logp, dt.type(10.0))
            x[:, d] = mu[:, d] + np.exp(m[:, d]) * u[:, d]
        ld_sum = ld_sum + m.sum(axis=1)
        u = x
    return u, ld_sum


def nde_sample_from_noise(spec, flat, std, u, y_obs):
    """jaxili/model.py:1116-1126 + 941-947 with the base noise ``u`` supplied:
    y'=(y-mean)/std broadcast to N rows ((C,) or (1,C) accepted), inverse flow,
    theta = scale*x + shift."""
    dt = flat.dtype
    y = np.asarray(y_obs, dtype=dt).reshape(-1, spec.n_cond)
    ys = (y - std.mean.astype(dt)) / std.std.astype(dt)
    ys = ys * np.ones((u.shape[0], 1), dtype=dt)
    x, _ = maf_inverse(spec, flat, u, ys)
    return std.scale.astype(dt) * x + std.shift.astype(dt)


# --------------------------------------------------------------------------
# A.4  loss and analytic gradients
# --------------------------------------------------------------------------
def nll_loss(spec, flat, std, x, y):
    """jaxili/loss.py:35-58 (NPE: x=theta, y=x) / 61-83 (NLE: arguments swapped
    by the caller): -mean(log_prob)."""
    return -nde_log_prob(spec, flat, std, x, y).mean()


def nll_loss_and_grad(spec, flat, std, x, y, inv_b=None, want_dy=False):
    """Value and gradient of the NLL w.r.t. the flat parameters, i.e. what
    ``jax.value_and_grad`` returns in jaxili/train.py:330-332, by hand-written
    reverse mode.  ``inv_b`` overrides 1/B for
    data-parallel shards (sum over shards of the returned values = global).
    With ``want_dy`` also returns d loss / d y' (standardised conditioner)."""
    dt = flat.dtype
    B = x.shape[0]
    inv_b = dt.type(1.0 / B if inv_b is None else inv_b)
    xs = (x.astype(dt) - std.shift.astype(dt)) * (dt.type(1.0) / std.scale.astype(dt))
    c = (-np.log(np.abs(std.scale.astype(dt)))).sum()
    ys = (y.astype(dt) - std.mean.astype(dt)) / std.std.astype(dt)
    u, ld, tape = maf_forward(spec, flat, xs, ys, keep=True)
    D = spec.n_in
    logp = dt.type(-0.5) * (u * u).sum(axis=1) - dt.type(0.5 * D * LOG_2PI) + ld + c
    loss = -(logp.sum()) * inv_b
    P = spec.unflatten(flat)
    grad = np.zeros_like(flat)
    G = spec.unflatten(grad)
    g_u = u * inv_b
    g_y = np.zeros_like(ys)
    n = spec.n_lin
    for k in reversed(range(spec.n_layers)):
        xin, hs, ds, s, v = tape[k]
        g_v = g_u[:, ::-1] if spec.use_reverse else g_u
        g_xdir = g_v * s
        g_a = np.concatenate([-g_v * s, dt.type(0.5) * g_v * v - dt.type(0.5) * inv_b], axis=1)
        for i in reversed(range(n)):
            W, _ = P[k][i]
            M = spec.masks[k][i].astype(dt)
            G[k][i][1][:] = g_a.sum(axis=0)
            G[k][i][0][:] = M * (hs[i].T @ g_a)
            g_h = g_a @ (M * W).T
            if i > 0:
                g_a = g_h * ds[i - 1]
        g_u = g_xdir + g_h[:, :D]
        if spec.n_cond:
            g_y = g_y + g_h[:, D:]
    if want_dy:
        return loss, grad, g_y
    return loss, grad


# --------------------------------------------------------------------------
# A.5  optimiser chain (optax semantics restated; jaxili/train.py:246-300)
# --------------------------------------------------------------------------
def warmup_cosine_lr(t, lr, warmup, decay_steps):
    """optax.warmup_cosine_decay_schedule(init=0.01*lr, peak=lr, warmup_steps,
    decay_steps, end=0.01*lr) evaluated at step count ``t`` (jaxili/train.py:278-284).
    = join_schedules([linear_schedule(init, peak, warmup),
                      cosine_decay_schedule(peak, decay_steps - warmup, alpha=end/peak)],
                     boundaries=[warmup])."""
    init = 0.01 * lr
    alpha = 0.01
    if t < warmup:
        # optax.linear_schedule: frac = 1 - count/transition_steps; (init-end)*frac + end
        if warmup <= 0:
            return lr
        frac = 1.0 - min(max(t, 0), warmup) / warmup
        return (init - lr) * frac + lr
    dec = decay_steps - warmup
    if dec <= 0:
        raise ValueError("cosine_decay_schedule requires positive decay_steps")
    tau = min(t - warmup, dec)
    cosine = 0.5 * (1.0 + math.cos(math.pi * tau / dec))
    return lr * ((1.0 - alpha) * cosine + alpha)


class AdamState:
    def __init__(self, n, dtype):
        self.m = np.zeros(n, dtype=dtype)
        self.v = np.zeros(n, dtype=dtype)
        self.count = 0


def clip_adam_step(flat, grad, st, lr, warmup, decay_steps, clip=5.0,
                   b1=0.9, b2=0.999, eps=1e-8):
    """optax.chain(clip_by_global_norm(clip), adam(schedule)) applied once
    (jaxili/train.py:286-292) followed by TrainState.apply_gradients
    (params += updates, train.py:333).  Returns (new_flat, lr_t, gnorm)."""
    dt = flat.dtype
    g = grad.astype(dt)
    gnorm = np.sqrt((g.astype(np.float64) ** 2).sum()) if dt == np.float64 else np.sqrt((g * g).sum(dtype=dt))
    # optax.clip_by_global_norm: trigger = g_norm < max_norm; g if trigger else g / g_norm * max_norm
    if not (gnorm < clip):
        g = (g / dt.type(gnorm)) * dt.type(clip)
    t = st.count
    st.m = dt.type(b1) * st.m + dt.type(1 - b1) * g
    st.v = dt.type(b2) * st.v + dt.type(1 - b2) * (g * g)
    tn = t + 1
    mhat = st.m / dt.type(1 - b1**tn)
    vhat = st.v / dt.type(1 - b2**tn)
    lr_t = warmup_cosine_lr(t, lr, warmup, decay_steps)
    upd = mhat / (np.sqrt(vhat) + dt.type(eps))
    st.count = tn
    return flat - dt.type(lr_t) * upd, lr_t, float(gnorm)
