"""CUDA path against the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from tests.util import grad_err, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*.npz"))))
def test_cuda_matches_golden(path):
    from jaxili_b200.engine import MafEngine
    z = np.load(path)
    dev = torch.device("cuda:0")
    eng = MafEngine(int(z["n_in"]), int(z["n_cond"]), int(z["n_layers"]), [int(v) for v in z["layers"]],
                    str(z["activation"]), bool(z["use_reverse"]), int(z["seed"]), z["shift"], z["scale"],
                    z["mean"], z["std"], device=dev)
    t = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)
    p, x, y = t(z["flat"]), t(z["x"]), t(z["y"])
    assert rel_err(eng.log_prob(p, x, y).cpu().numpy(), z["logp"]) < 1e-5          # FP32 tolerance: 1e-5
    loss, grad = torch.zeros(1, device=dev), torch.zeros_like(p)
    eng.loss_grad(p, x, y, loss, grad)
    assert rel_err(loss.item(), float(z["loss"])) < 1e-5
    assert grad_err(grad.cpu().numpy(), z["grad"]) < 1e-5
    # conditioner gradient (NLE / HMC entry point); a ReLU pre-activation within FP32 rounding of 0 may take the
    # other branch than the float64 fixture in at most one row
    lp2, dy = eng.log_prob_grad_y(p, x, y)
    assert rel_err(lp2.cpu().numpy(), z["logp"]) < 1e-5
    row_err = np.abs(dy.cpu().numpy() - z["dlogp_dy"]).max(axis=1) / np.abs(z["dlogp_dy"]).max()
    assert int((row_err >= 1e-5).sum()) <= (1 if str(z["activation"]) == "relu" else 0), row_err.max()
    s = eng.sample_from_noise(p, t(z["u"]), y[:1])
    assert rel_err(s.cpu().numpy(), z["samples"]) < 2e-5
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    eng.clip_adam(p, grad, m, v, step, eng.make_opt_desc(lr=5e-4, warmup=0.1, decay_steps=1e6, clip=5.0),
                  norm_current=True)
    assert np.max(np.abs(p.cpu().numpy() - z["params_after_adam"])) < 1e-6


def test_full_size_properties():
    """BASELINE.json full sizes through size-independent properties: the batch-65536 loss equals the
    mean of shard losses, gradients are linear in the shard decomposition, sampling round-trips."""
    from jaxili_b200.engine import MafEngine
    from oracle import maf_numpy as onp
    from tests.util import CONFIGS, make_case
    spec, flat, std, x, y = make_case(CONFIGS["c3"], 16384, seed=21)
    dev = torch.device("cuda:0")
    eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, "relu", True, 42, std.shift, std.scale,
                    std.mean, std.std, device=dev)
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
    p, xt, yt = t(flat), t(x), t(y)
    loss, grad = torch.zeros(1, device=dev), torch.zeros_like(p)
    eng.loss_grad(p, xt, yt, loss, grad)
    lp = eng.log_prob(p, xt, yt)
    assert abs(loss.item() + lp.double().mean().item()) < 1e-5 * max(1, abs(loss.item()))
    acc = torch.zeros_like(p, dtype=torch.float64)
    l_acc = 0.0
    g2, l2 = torch.zeros_like(p), torch.zeros(1, device=dev)
    for s in (slice(0, 5000), slice(5000, 5004), slice(5004, 16384)):
        eng.loss_grad(p, xt[s].contiguous(), yt[s].contiguous(), l2, g2, inv_global_b=1 / 16384, packed_current=True)
        acc += g2.double()
        l_acc += l2.item()
    assert abs(l_acc - loss.item()) < 1e-5 * max(1, abs(loss.item()))
    assert float((acc - grad.double()).abs().max() / grad.abs().max()) < 1e-5
    # 1M-row sampling: encode -> decode round trip through the two kernels
    N = 1 << 20
    th = eng.sample_philox(p, N, yt[:64], seed=3, rows_per_obs=N // 64)
    u, ld = eng.forward(p, th, yt[:64].repeat_interleave(N // 64, dim=0))
    assert torch.isfinite(th).all()
    # base-space points are standard normal draws: mean ~ 0, var ~ 1
    assert abs(u.mean().item()) < 5e-3 and abs(u.var().item() - 1) < 5e-3
