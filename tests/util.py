"""Shared helpers for the parity tests (oracle side + seeded inputs)."""
import numpy as np

from oracle import maf_numpy as onp

# BASELINE.json configs: (D, C, L, hidden)
CONFIGS = {
    "c1": dict(n_in=2, n_cond=2, n_layers=5, layers=[50, 50]),
    "c2": dict(n_in=10, n_cond=10, n_layers=5, layers=[50, 50]),
    "c3": dict(n_in=8, n_cond=5, n_layers=5, layers=[50, 50]),
    "ref_test": dict(n_in=3, n_cond=5, n_layers=3, layers=[128, 128]),  # jaxili/tests/test_nn.py:10-24
}


def make_case(cfg, B, seed=0, activation="relu", use_reverse=True, scale=0.12, standardize=True):
    """Seeded (spec, flat params, standardisation, x, y) at 'trained-scale' weights so that
    exp / clamp paths are exercised."""
    rng = np.random.default_rng(seed)
    spec = onp.MafSpec(activation=activation, use_reverse=use_reverse, seed=42, **cfg)
    flat = spec.init_params(rng, scale=scale, dtype=np.float32)
    # biases non-zero too
    flat += (rng.standard_normal(flat.size) * 0.02).astype(np.float32) * (spec.flat_mask() > 0)
    D, C = spec.n_in, spec.n_cond
    if standardize:
        std = onp.Standardization(rng.normal(size=D).astype(np.float32),
                                  rng.uniform(0.5, 2.0, size=D).astype(np.float32),
                                  rng.normal(size=C).astype(np.float32),
                                  rng.uniform(0.5, 2.0, size=C).astype(np.float32))
    else:
        std = onp.Standardization.identity(D, C)
    x = (rng.standard_normal((B, D)) * std.scale + std.shift).astype(np.float32)
    y = (rng.standard_normal((B, C)) * std.std + std.mean).astype(np.float32)
    return spec, flat, std, x, y


def rel_err(a, b):
    """max |a-b| / max(1, |b|): log_prob crosses zero, so a pure relative error is ill-posed."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def grad_err(g, g_ref):
    """max |g-g_ref| / max|g_ref| (gradient entries span many orders of magnitude)."""
    g = np.asarray(g, dtype=np.float64)
    g_ref = np.asarray(g_ref, dtype=np.float64)
    return float(np.max(np.abs(g - g_ref)) / max(np.max(np.abs(g_ref)), 1e-30))
