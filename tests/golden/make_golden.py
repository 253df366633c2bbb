"""Generates the committed golden fixtures under tests/golden/ from the CPU oracle.

    python tests/golden/make_golden.py

PARITY UNPINNED: jaxili itself (JAX/flax/optax/distrax) cannot be imported in this image, so these
vectors are outputs of the line-by-line restatement in oracle/maf_numpy.py (FP64), not of the
reference.  What they pin: (1) the oracle against accidental edits, (2) the CUDA path at fixed,
reviewable inputs.  Known answers that ARE derivable without JAX are stored too: the per-layer
seeds and MADE degrees from NumPy's legacy global RNG (jaxili/model.py:583,613-616,831-843).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import maf_numpy as onp  # noqa: E402
from tests.util import CONFIGS, make_case  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [  # name, config, B, activation, reverse
    ("c1_relu", "c1", 128, "relu", True),
    ("c2_relu", "c2", 96, "relu", True),
    ("c3_relu_norev", "c3", 64, "relu", False),
    ("ref_test_silu", "ref_test", 10, "silu", True),
]


def main():
    for tag, cname, B, act, rev in CASES:
        spec, flat, std, x, y = make_case(CONFIGS[cname], B, seed=123, activation=act, use_reverse=rev)
        f64 = flat.astype(np.float64)
        x64, y64 = x.astype(np.float64), y.astype(np.float64)
        logp = onp.nde_log_prob(spec, f64, std, x64, y64)
        loss, grad = onp.nll_loss_and_grad(spec, f64, std, x64, y64)
        u = np.random.default_rng(5).standard_normal((32, spec.n_in)).astype(np.float32)
        samples = onp.nde_sample_from_noise(spec, f64, std, u.astype(np.float64), y[0])
        # per-row gradient of log_prob w.r.t. the RAW conditioning input (the NLE / HMC entry point): the reverse
        # pass with 1/B := -1 differentiates sum_b log p_b; d/dy = (d/dy') / std
        _, _, gy = onp.nll_loss_and_grad(spec, f64, std, x64, y64, inv_b=-1.0, want_dy=True)
        dlogp_dy = gy / std.std.astype(np.float64)
        st = onp.AdamState(f64.size, np.float64)
        p1, lr0, gn = onp.clip_adam_step(f64, grad, st, 5e-4, 0.1, 1e6, clip=5.0)
        np.savez_compressed(
            os.path.join(HERE, f"{tag}.npz"),
            n_in=spec.n_in, n_cond=spec.n_cond, n_layers=spec.n_layers, layers=np.array(spec.layers),
            activation=act, use_reverse=rev, seed=42, layer_seeds=np.array(spec.seeds),
            flat=flat, x=x, y=y, shift=std.shift, scale=std.scale, mean=std.mean, std=std.std,
            logp=logp, loss=loss, grad=grad.astype(np.float32), u=u, samples=samples,
            params_after_adam=p1.astype(np.float32), lr0=lr0, gnorm=gn, dlogp_dy=dlogp_dy)
        print(tag, "P =", spec.n_params, "loss =", float(loss))


if __name__ == "__main__":
    main()
