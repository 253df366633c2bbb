"""CPU tests of the oracle itself: what can be pinned without JAX is pinned here."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import maf_numpy as onp
from oracle.maf_torch import TorchMaf, TorchTrainer
from tests.util import CONFIGS, make_case, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_layer_seeds_known_answer():
    """SURVEY F7 (scratch-verified against NumPy's legacy RNG): seed 42 -> these MAFLayer seeds."""
    assert onp.layer_seeds(42, 10) == [102, 435, 860, 270, 106, 71, 700, 20, 614, 121]


def test_mask_structure():
    """jaxili/model.py:603-663: shapes, conditioning rows all ones, strict output mask duplicated,
    autoregressive property of the mask product."""
    for cfg in CONFIGS.values():
        spec = onp.MafSpec(seed=42, **cfg)
        D, C = spec.n_in, spec.n_cond
        for ms in spec.masks:
            assert ms[0].shape == (D + C, spec.layers[0]) and (ms[0][D:] == 1).all()
            assert ms[-1].shape == (spec.layers[-1], 2 * D)
            assert (ms[-1][:, :D] == ms[-1][:, D:]).all()
            conn = ms[0][:D]
            for m in ms[1:]:
                conn = conn @ m
            conn = conn[:, :D]            # paths from x_i to mu_j
            assert (conn[np.tril_indices(D)] == 0).all()   # out_j depends only on x_i with i < j
    # n_in = 2: all hidden degrees are 0; n_in = 1 raises like the reference
    spec2 = onp.MafSpec(2, 2, 1, [8], seed=1)
    assert (onp.made_degrees(2, [8], 5)[1] == 0).all()
    with pytest.raises(ValueError):
        onp.MafSpec(1, 2, 1, [8], seed=1)


@pytest.mark.parametrize("act", ["relu", "silu", "tanh", "gelu", "elu", "sigmoid", "softplus", "leaky_relu"])
def test_analytic_gradient_vs_autograd_and_fd(act):
    spec, flat, std, x, y = make_case(CONFIGS["c3"], 24, seed=3, activation=act)
    f64, x64, y64 = flat.astype(np.float64), x.astype(np.float64), y.astype(np.float64)
    loss, grad, gy = onp.nll_loss_and_grad(spec, f64, std, x64, y64, want_dy=True)
    l2, g2, gy2 = TorchMaf(spec, std, torch.float64).loss_and_grad(f64, x64, y64, want_dy=True)
    assert abs(loss - l2) < 1e-12 * max(1, abs(l2))
    assert np.max(np.abs(grad - g2)) < 1e-10 * np.max(np.abs(g2))
    # d loss / d y (raw conditioner) = d/dy' / std
    assert np.max(np.abs(gy / std.std - gy2)) < 1e-10 * max(np.max(np.abs(gy2)), 1e-30)
    assert (grad[spec.flat_mask() == 0] == 0).all()
    rng = np.random.default_rng(0)
    nz = np.flatnonzero(spec.flat_mask())
    for i in rng.choice(nz, 6, replace=False):
        e = np.zeros_like(f64)
        e[i] = 1e-6
        fd = (onp.nll_loss(spec, f64 + e, std, x64, y64) - onp.nll_loss(spec, f64 - e, std, x64, y64)) / 2e-6
        assert abs(fd - grad[i]) < 1e-6 * max(1.0, abs(grad[i]))


def test_round_trip_invariant():
    """jaxili/tests/test_nn.py:35-41: forward o backward = id and log_det = -log_det_inverse (1e-5),
    here in the FP32 oracle, for both use_reverse settings."""
    for rev, L in ((True, 3), (False, 4)):
        spec = onp.MafSpec(3, 5, L, [128, 128], activation="silu", use_reverse=rev, seed=42)
        rng = np.random.default_rng(1)
        flat = spec.init_params(rng, 0.05, np.float32)
        x = rng.standard_normal((10, 3)).astype(np.float32)
        y = rng.standard_normal((10, 5)).astype(np.float32)
        u, ld = onp.maf_forward(spec, flat, x, y)
        xr, ldr = onp.maf_inverse(spec, flat, u, y)
        np.testing.assert_allclose(x, xr, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(ld, -ldr, rtol=1e-5, atol=1e-5)


def test_inverse_clamp_breaks_round_trip_by_design():
    """the inverse clamps -logp/2 at 10, the forward does not."""
    spec = onp.MafSpec(2, 0, 1, [4], seed=3)
    flat = np.zeros(spec.n_params)
    lay = spec.unflatten(flat)
    lay[0][-1][1][2:] = -30.0          # bias of logp: -0.5*logp = 15 > 10
    u = np.ones((1, 2))
    x, ld = onp.maf_inverse(spec, flat, u, np.zeros((1, 0)))
    assert np.allclose(x, np.exp(10.0)) and np.allclose(ld, 20.0)


def test_schedule_and_adam_vs_torch():
    """warmup-cosine (optax formula) known points; the Adam update against torch.optim.Adam."""
    lr = 5e-4
    assert onp.warmup_cosine_lr(0, lr, 0.1, 1000) == pytest.approx(0.01 * lr)
    assert onp.warmup_cosine_lr(1, lr, 0.1, 1000) == pytest.approx(lr, rel=1e-5)
    assert onp.warmup_cosine_lr(10**9, lr, 0.1, 1000) == pytest.approx(0.01 * lr)
    assert onp.warmup_cosine_lr(5, lr, 10, 1000) == pytest.approx(0.01 * lr + 0.99 * lr * 0.5)
    mid = onp.warmup_cosine_lr(10 + 495, lr, 10, 1000)
    assert mid == pytest.approx(lr * (0.99 * 0.5 + 0.01))
    rng = np.random.default_rng(0)
    p = rng.standard_normal(50)
    pt = torch.tensor(p.copy(), requires_grad=True)
    opt = torch.optim.Adam([pt], lr=1.0, betas=(0.9, 0.999), eps=1e-8)
    st = onp.AdamState(50, np.float64)
    for t in range(5):
        g = rng.standard_normal(50) * 0.01           # norm < clip: no clipping
        lr_t = onp.warmup_cosine_lr(t, 1e-3, 0.1, 100)
        for grp in opt.param_groups:
            grp["lr"] = lr_t
        pt.grad = torch.tensor(g)
        opt.step()
        p, lr_o, gn = onp.clip_adam_step(p, g, st, 1e-3, 0.1, 100)
        assert lr_o == lr_t
    assert np.max(np.abs(p - pt.detach().numpy())) < 1e-12
    # clipping: norm 10 with clip 5 halves the gradient
    st2 = onp.AdamState(4, np.float64)
    g = np.array([6.0, 8.0, 0.0, 0.0])
    p2, _, gn = onp.clip_adam_step(np.zeros(4), g, st2, 1.0, 0.0, 10, clip=5.0)
    assert gn == pytest.approx(10.0) and np.allclose(st2.m, 0.1 * g / 2)


def test_torch_trainer_matches_numpy_chain():
    spec, flat, std, x, y = make_case(CONFIGS["c1"], 64, seed=2)
    tr = TorchTrainer(TorchMaf(spec, std, torch.float64), flat.astype(np.float64), lr=1e-2, warmup=2, decay_steps=20, clip=0.5)
    ref = flat.astype(np.float64)
    st = onp.AdamState(ref.size, np.float64)
    for _ in range(4):
        tr.train_step(torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64))
        _, g = onp.nll_loss_and_grad(spec, ref, std, x.astype(np.float64), y.astype(np.float64))
        ref, _, _ = onp.clip_adam_step(ref, g, st, 1e-2, 2, 20, clip=0.5)
    assert np.max(np.abs(tr.flat.detach().numpy() - ref)) < 1e-12


def test_data_parallel_shards_sum_to_global():
    """shard losses/gradients computed with 1/B_global add up to the global ones."""
    spec, flat, std, x, y = make_case(CONFIGS["c2"], 64, seed=4)
    f64 = flat.astype(np.float64)
    L, G = onp.nll_loss_and_grad(spec, f64, std, x.astype(np.float64), y.astype(np.float64))
    parts = [onp.nll_loss_and_grad(spec, f64, std, x[s].astype(np.float64), y[s].astype(np.float64), inv_b=1 / 64)
             for s in (slice(0, 16), slice(16, 64))]
    assert abs(sum(p[0] for p in parts) - L) < 1e-12 * abs(L)
    assert np.max(np.abs(sum(p[1] for p in parts) - G)) < 1e-12 * np.max(np.abs(G))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*.npz"))))
def test_oracle_reproduces_golden(path):
    z = np.load(path)
    spec = onp.MafSpec(int(z["n_in"]), int(z["n_cond"]), int(z["n_layers"]), list(z["layers"]),
                       activation=str(z["activation"]), use_reverse=bool(z["use_reverse"]), seed=int(z["seed"]))
    assert spec.seeds == list(z["layer_seeds"])
    std = onp.Standardization(z["shift"], z["scale"], z["mean"], z["std"])
    f64 = z["flat"].astype(np.float64)
    lp = onp.nde_log_prob(spec, f64, std, z["x"].astype(np.float64), z["y"].astype(np.float64))
    assert rel_err(lp, z["logp"]) < 1e-12
    # FP32 evaluation of the same model sits inside the tolerance the kernels are held to
    lp32 = onp.nde_log_prob(spec, z["flat"], std, z["x"], z["y"])
    assert rel_err(lp32, z["logp"]) < 5e-6
    loss, grad = onp.nll_loss_and_grad(spec, f64, std, z["x"].astype(np.float64), z["y"].astype(np.float64))
    assert abs(loss - float(z["loss"])) < 1e-12 * max(1, abs(loss))
    assert np.max(np.abs(grad - z["grad"])) < 1e-6 * np.max(np.abs(grad))
    s = onp.nde_sample_from_noise(spec, f64, std, z["u"].astype(np.float64), z["y"][0])
    assert rel_err(s, z["samples"]) < 1e-12
    ts = TorchMaf(spec, std, torch.float64).sample_from_noise(f64, z["u"].astype(np.float64), z["y"][0])
    assert rel_err(ts, z["samples"]) < 1e-10
    # gradient of log_prob w.r.t. the raw conditioning input: hand-written reverse pass == stored == torch autograd
    _, _, gy = onp.nll_loss_and_grad(spec, f64, std, z["x"].astype(np.float64), z["y"].astype(np.float64),
                                     inv_b=-1.0, want_dy=True)
    assert np.max(np.abs(gy / z["std"].astype(np.float64) - z["dlogp_dy"])) < 1e-12 * max(1.0, np.abs(z["dlogp_dy"]).max())
    yt = torch.tensor(z["y"].astype(np.float64), requires_grad=True)
    TorchMaf(spec, std, torch.float64).log_prob(torch.tensor(f64), torch.tensor(z["x"].astype(np.float64)), yt).sum().backward()
    assert np.max(np.abs(yt.grad.numpy() - z["dlogp_dy"])) < 1e-9 * max(1.0, np.abs(z["dlogp_dy"]).max())
