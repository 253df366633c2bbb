"""MADE degree / mask construction on the host (product side).

Follows jaxili/model.py:583, 603-663 (``ConditionalMADE._create_masks``) and
829-843 (per-layer seeds in ``ConditionalMAF.setup``): the legacy global NumPy RNG is
seeded and consumed in exactly the reference's order, so the same ``seed`` gives the
same masks as jaxili (seed 42 -> layer seeds 102, 435, 860, 270, 106, ...).

Only the hidden-level degrees are kept: the native library derives the 0/1 masks from
them (``mafb200_create``), and the degree-sorted sampler needs them anyway.
"""

from __future__ import annotations

import numpy as np


def maf_layer_seeds(seed, n_layers):
    """One ``np.random.randint(0, 1000)`` per MAF layer after ``np.random.seed(seed)``."""
    state = np.random.get_state()
    try:
        np.random.seed(seed)
        return [int(np.random.randint(0, 1000)) for _ in range(n_layers)]
    finally:
        np.random.set_state(state)


def made_hidden_degrees(n_in, hidden_dims, seed):
    """Hidden degree vectors m^1..m^n of one MADE (natural input order m^0 = arange)."""
    if n_in < 2:
        raise ValueError("n_in must be >= 2: jaxili draws randint(low, n_in - 1)")
    state = np.random.get_state()
    try:
        np.random.seed(seed)
        prev_min = 0  # min(arange(n_in))
        out = []
        for h in hidden_dims:
            m = np.random.randint(prev_min, n_in - 1, size=h)
            out.append(m.astype(np.int32))
            prev_min = int(m.min())
        return out
    finally:
        np.random.set_state(state)


def maf_degrees(n_in, hidden_dims, n_layers, seed):
    """int32 array (n_layers, sum(hidden_dims)): concatenated hidden degrees per layer."""
    seeds = maf_layer_seeds(seed, n_layers)
    rows = [np.concatenate(made_hidden_degrees(n_in, hidden_dims, s)) for s in seeds]
    return np.ascontiguousarray(np.stack(rows).astype(np.int32))


def masks_from_degrees(n_in, n_cond, hidden_dims, degrees_row):
    """0/1 (in, out) matrices of one MADE from its concatenated hidden degrees
    (host mirror of what the library builds; used by tests and parameter I/O)."""
    D = n_in
    deg = [np.arange(D)]
    off = 0
    for h in hidden_dims:
        deg.append(np.asarray(degrees_row[off:off + h]))
        off += h
    deg.append(deg[0])
    mats = []
    for i in range(len(deg) - 1):
        a, b = deg[i], deg[i + 1]
        if i == len(deg) - 2:
            M = (a[:, None] < b[None, :]).astype(np.uint8)
            M = np.concatenate([M, M], axis=1)
        else:
            M = (a[:, None] <= b[None, :]).astype(np.uint8)
            if i == 0:
                M = np.concatenate([M, np.ones((n_cond, len(b)), dtype=np.uint8)], axis=0)
        mats.append(M)
    return mats
