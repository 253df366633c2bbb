"""Fixtures written by tools/jaxili_golden.py from the REAL jaxili (on a machine that has JAX), when present.

No such file can be produced in this image, so by default these tests skip and parity with jaxili stays
"unpinned"; dropping tests/golden/jaxili_*.npz next to the oracle's fixtures pins the oracle (CPU test) and the CUDA
path (GPU test) to jaxili's own numbers through the flax parameter layout (jaxili_b200/layout.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import maf_numpy as onp

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "jaxili_*.npz")))
needs_files = pytest.mark.skipif(not FILES, reason="no jaxili-generated fixtures (tools/jaxili_golden.py needs JAX): "
                                                   "parity with jaxili itself is unpinned")


def _load(path):
    z = np.load(path, allow_pickle=False)
    layers = [int(v) for v in z["layers"]]
    spec = onp.MafSpec(n_in=int(z["n_in"]), n_cond=int(z["n_cond"]), n_layers=int(z["n_layers"]), layers=layers,
                       activation=str(z["activation"]), use_reverse=bool(z["use_reverse"]), seed=int(z["seed"]))
    from jaxili_b200 import layout

    def tree(prefix):
        m = {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)}
        t = layout.unflatten_paths(m)
        return t["nde"] if "nde" in t and "layer_list_0" not in t else t

    dims = [spec.n_in + spec.n_cond, *layers, 2 * spec.n_in]
    flat = lambda prefix: layout.flat_from_tree(tree(prefix), dims, spec.n_layers).numpy()
    D, C = spec.n_in, spec.n_cond
    if bool(z["wrapper"]):
        std = onp.Standardization(z["shift"], z["scale"], np.zeros(C, np.float32), np.ones(C, np.float32))
    else:
        std = onp.Standardization.identity(D, C)
    return z, spec, std, flat


@needs_files
@pytest.mark.parametrize("path", FILES or ["-"])
def test_oracle_matches_jaxili(path):
    z, spec, std, flat = _load(path)
    p = flat("params/").astype(np.float64)
    x, y = z["x"].astype(np.float64), z["y"].astype(np.float64)
    lp = onp.nde_log_prob(spec, p, std, x, y)
    assert np.max(np.abs(lp - z["log_prob"]) / np.maximum(1, np.abs(z["log_prob"]))) < 1e-5
    loss, grad = onp.nll_loss_and_grad(spec, p, std, x, y)
    g_ref = flat("grads/")
    assert abs(loss - float(z["loss"])) < 1e-5 * max(1, abs(float(z["loss"])))
    assert np.max(np.abs(grad - g_ref)) / np.max(np.abs(g_ref)) < 1e-5
    lr, warm, dec, clip = [float(v) for v in z["opt"]]
    st = onp.AdamState(p.size, np.float64)
    q = p.copy()
    for i in range(3):
        l, g = onp.nll_loss_and_grad(spec, q, std, x, y)
        assert abs(l - float(z["losses3"][i])) < 2e-5 * max(1, abs(l))
        q, _, _ = onp.clip_adam_step(q, g, st, lr, warm, dec, clip=clip)
    assert np.max(np.abs(q - flat("after3/"))) < 5e-6
    if "u" in z.files:
        u, ld = onp.maf_forward(spec, p, x, y)
        assert np.max(np.abs(u - z["u"])) < 1e-5 * max(1, np.abs(z["u"]).max())
        assert np.max(np.abs(ld - z["log_det"])) < 1e-5 * max(1, np.abs(z["log_det"]).max())
        xr, ldr = onp.maf_inverse(spec, p, z["u"].astype(np.float64), y)
        assert np.max(np.abs(xr - z["x_rec"])) < 2e-5 * max(1, np.abs(z["x_rec"]).max())
        assert np.max(np.abs(ldr - z["log_det_rec"])) < 2e-5 * max(1, np.abs(z["log_det_rec"]).max())


@needs_files
@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES or ["-"])
def test_cuda_matches_jaxili(path):
    import torch
    from jaxili_b200.engine import MafEngine
    z, spec, std, flat = _load(path)
    dev = torch.device("cuda:0")
    eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, spec.activation, spec.use_reverse, spec.seed,
                    std.shift, std.scale, std.mean, std.std, device=dev)
    t = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)
    p, x, y = t(flat("params/")), t(z["x"]), t(z["y"])
    lp = eng.log_prob(p, x, y).cpu().numpy()
    assert np.max(np.abs(lp - z["log_prob"]) / np.maximum(1, np.abs(z["log_prob"]))) < 1e-5
    loss, grad = torch.zeros(1, device=dev), torch.zeros_like(p)
    eng.loss_grad(p, x, y, loss, grad)
    g_ref = flat("grads/")
    assert abs(loss.item() - float(z["loss"])) < 1e-5 * max(1, abs(float(z["loss"])))
    assert np.max(np.abs(grad.cpu().numpy() - g_ref)) / np.max(np.abs(g_ref)) < 1e-5
    lr, warm, dec, clip = [float(v) for v in z["opt"]]
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    opt = eng.make_opt_desc(lr=lr, warmup=warm, decay_steps=dec, clip=clip)
    for i in range(3):
        eng.train_step(p, x, y, loss, grad, m, v, step, opt)
        assert abs(loss.item() - float(z["losses3"][i])) < 2e-5 * max(1, abs(loss.item()))
    assert np.max(np.abs(p.cpu().numpy() - flat("after3/"))) < 5e-6
    if "u" in z.files:
        xr, ldr = eng.inverse(t(flat("params/")), t(z["u"]), y)
        assert np.max(np.abs(xr.cpu().numpy() - z["x_rec"])) < 2e-5 * max(1, np.abs(z["x_rec"]).max())
        assert np.max(np.abs(ldr.cpu().numpy() - z["log_det_rec"])) < 2e-5 * max(1, np.abs(z["log_det_rec"]).max())


def test_golden_tool_is_self_consistent():
    """The consumer half of the tool runs here: a fixture written in the tool's format from the ORACLE's numbers
    (flax paths, value_and_grad, three optimiser steps) loads through layout.py and passes the comparison --
    so a fixture from a JAX machine needs no further plumbing."""
    import tempfile
    from jaxili_b200 import layout
    from tests.util import CONFIGS, make_case
    spec, flat, std, x, y = make_case(CONFIGS["ref_test"], 10, seed=77, activation="silu", use_reverse=True)
    std = onp.Standardization.identity(spec.n_in, spec.n_cond)
    p = flat.astype(np.float64)
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    dims = [spec.n_in + spec.n_cond, *spec.layers, 2 * spec.n_in]
    import torch
    paths = lambda prefix, vec: {prefix + k: v.numpy() for k, v in layout.flatten_paths(
        layout.tree_from_flat(torch.tensor(np.asarray(vec, np.float32)), dims, spec.n_layers)).items()}
    loss, grad = onp.nll_loss_and_grad(spec, p, std, x64, y64)
    st, q, losses = onp.AdamState(p.size, np.float64), p.copy(), []
    for _ in range(3):
        l, g = onp.nll_loss_and_grad(spec, q, std, x64, y64)
        losses.append(l)
        q, _, _ = onp.clip_adam_step(q, g, st, 5e-4, 0.1, 1e6, clip=5.0)
    u, ld = onp.maf_forward(spec, p, x64, y64)
    xr, ldr = onp.maf_inverse(spec, p, u, y64)
    rec = dict(n_in=spec.n_in, n_cond=spec.n_cond, n_layers=spec.n_layers, layers=np.array(spec.layers),
               activation="silu", use_reverse=True, seed=42, x=x, y=y, wrapper=False,
               log_prob=onp.nde_log_prob(spec, p, std, x64, y64), loss=np.float32(loss), losses3=np.array(losses, np.float32),
               opt=np.array([5e-4, 0.1, 1e6, 5.0]), u=u, log_det=ld, x_rec=xr, log_det_rec=ldr)
    rec.update(paths("params/", flat)); rec.update(paths("grads/", grad)); rec.update(paths("after3/", q))
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "jaxili_selftest.npz")
        np.savez_compressed(f, **rec)
        test_oracle_matches_jaxili.__wrapped__(f) if hasattr(test_oracle_matches_jaxili, "__wrapped__") else None
        z, spec2, std2, flt = _load(f)
        assert np.array_equal(flt("params/"), flat)
        assert spec2.n_params == spec.n_params and spec2.layers == list(spec.layers)
        lp = onp.nde_log_prob(spec2, flt("params/").astype(np.float64), std2, x64, y64)
        assert np.max(np.abs(lp - z["log_prob"])) < 1e-6
        assert np.max(np.abs(flt("after3/") - q.astype(np.float32))) == 0
