"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (north_star): log_prob, loss, gradients within 1e-5 relative in FP32 -- measured as
|a-b| / max(1,|b|) for log_prob/loss (they cross zero) and max|dg| / max|g| for gradients -- against
the FP64 oracle; samples match for the same base noise.
"""
import numpy as np
import pytest
import torch

from oracle import maf_numpy as onp
from tests.util import CONFIGS, grad_err, make_case, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _engine(spec, std, device="cuda:0", flow="auto"):
    from jaxili_b200.engine import MafEngine
    eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, spec.activation,
                    spec.use_reverse, spec.seed, std.shift, std.scale, std.mean, std.std,
                    device=torch.device(device))
    eng.set_flow_kernel(flow)   # both fused kernels (barrier-phased, chain) serve every test they apply to
    return eng


FLOWS = ["phased", "chain"]


def _t(a):
    return torch.tensor(np.ascontiguousarray(a), device="cuda:0")


@pytest.mark.parametrize("name,B", [("c1", 128), ("c2", 4096), ("c3", 1000), ("c2", 7), ("c2", 33),
                                    ("ref_test", 10), ("c3", 16384)])
@pytest.mark.parametrize("flow", FLOWS)
def test_log_prob_matches_oracle(name, B, flow):
    spec, flat, std, x, y = make_case(CONFIGS[name], B, seed=1)
    eng = _engine(spec, std, flow=flow)
    assert (eng.mask() == spec.flat_mask().astype(np.uint8)).all()
    lp = eng.log_prob(_t(flat), _t(x), _t(y)).cpu().numpy()
    ref = onp.nde_log_prob(spec, flat.astype(np.float64), std, x.astype(np.float64), y.astype(np.float64))
    assert np.isfinite(lp).all()
    assert rel_err(lp, ref) < TOL


@pytest.mark.parametrize("act", ["silu", "tanh", "gelu", "elu", "sigmoid", "softplus", "leaky_relu"])
@pytest.mark.parametrize("flow", FLOWS)
def test_log_prob_and_grad_activations(act, flow):
    spec, flat, std, x, y = make_case(CONFIGS["ref_test"], 50, seed=2, activation=act, use_reverse=False)
    eng = _engine(spec, std, flow=flow)
    lp = eng.log_prob(_t(flat), _t(x), _t(y)).cpu().numpy()
    ref = onp.nde_log_prob(spec, flat.astype(np.float64), std, x.astype(np.float64), y.astype(np.float64))
    assert rel_err(lp, ref) < TOL
    loss = torch.zeros(1, device="cuda:0")
    grad = torch.zeros(spec.n_params, device="cuda:0")
    eng.loss_grad(_t(flat), _t(x), _t(y), loss, grad)
    l_ref, g_ref = onp.nll_loss_and_grad(spec, flat.astype(np.float64), std, x.astype(np.float64), y.astype(np.float64))
    assert rel_err(loss.item(), l_ref) < TOL
    assert grad_err(grad.cpu().numpy(), g_ref) < TOL


@pytest.mark.parametrize("name,B", [("c1", 128), ("c2", 4096), ("c3", 16384), ("c2", 130), ("c2", 5),
                                    ("ref_test", 64)])
@pytest.mark.parametrize("flow", FLOWS)
def test_loss_and_grad_match_oracle(name, B, flow):
    spec, flat, std, x, y = make_case(CONFIGS[name], B, seed=3)
    eng = _engine(spec, std, flow=flow)
    loss = torch.zeros(1, device="cuda:0")
    grad = torch.full((spec.n_params,), 7.0, device="cuda:0")
    eng.loss_grad(_t(flat), _t(x), _t(y), loss, grad)
    l_ref, g_ref = onp.nll_loss_and_grad(spec, flat.astype(np.float64), std, x.astype(np.float64), y.astype(np.float64))
    g = grad.cpu().numpy()
    assert rel_err(loss.item(), l_ref) < TOL
    assert grad_err(g, g_ref) < TOL
    # masked weights get exactly zero gradient (they keep their init value forever)
    assert (g[spec.flat_mask() == 0] == 0).all()
    # deterministic two-stage reduction: bitwise reproducible
    grad2 = torch.zeros_like(grad)
    eng.loss_grad(_t(flat), _t(x), _t(y), loss, grad2)
    assert torch.equal(grad, grad2)


def test_device_side_batch_selection():
    spec, flat, std, x, y = make_case(CONFIGS["c2"], 4 * 256, seed=4)
    eng = _engine(spec, std)
    p, xt, yt = _t(flat), _t(x), _t(y)
    loss = torch.zeros(1, device="cuda:0")
    g_sel = torch.zeros(spec.n_params, device="cuda:0")
    g_ref = torch.zeros(spec.n_params, device="cuda:0")
    for idx in (0, 1, 3, 6):
        bi = torch.tensor([idx], dtype=torch.int32, device="cuda:0")
        eng.loss_grad(p, xt, yt, loss, g_sel, batch_rows=256, batch_index=bi, n_batches=4)
        l1 = loss.item()
        b = idx % 4
        eng.loss_grad(p, xt[b * 256:(b + 1) * 256].contiguous(), yt[b * 256:(b + 1) * 256].contiguous(), loss, g_ref)
        assert l1 == loss.item()
        assert torch.equal(g_sel, g_ref)


@pytest.mark.parametrize("kind", ["adam", "adamw", "sgd"])
def test_clip_adam_matches_oracle(kind):
    spec, flat, std, x, y = make_case(CONFIGS["c2"], 512, seed=5)
    eng = _engine(spec, std)
    p = _t(flat)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int32, device="cuda:0")
    loss = torch.zeros(1, device="cuda:0")
    grad = torch.zeros_like(p)
    hp = dict(lr=5e-3, warmup=0.1, decay_steps=50.0, clip=0.5)  # clip small enough to trigger
    opt = eng.make_opt_desc(kind=kind, weight_decay=1e-4 if kind != "adam" else 0.0, **hp)
    ref = flat.astype(np.float64)
    st = onp.AdamState(ref.size, np.float64)
    xt, yt = _t(x), _t(y)
    for it in range(6):
        eng.loss_grad(p, xt, yt, loss, grad, packed_current=it > 0)
        eng.clip_adam(p, grad, m, v, step, opt, norm_current=True)
        l_ref, g_ref = onp.nll_loss_and_grad(spec, ref, std, x.astype(np.float64), y.astype(np.float64))
        assert rel_err(loss.item(), l_ref) < 5 * TOL
        if kind == "adam":
            ref, lr_t, gn = onp.clip_adam_step(ref, g_ref, st, hp["lr"], hp["warmup"], hp["decay_steps"], clip=hp["clip"])
        else:
            gn = np.sqrt((g_ref ** 2).sum())
            g = g_ref if gn < hp["clip"] else g_ref / gn * hp["clip"]
            lr_t = onp.warmup_cosine_lr(st.count, hp["lr"], hp["warmup"], hp["decay_steps"])
            if kind == "sgd":
                ref = ref - lr_t * (g + 1e-4 * ref)
                st.count += 1
            else:
                st.m = 0.9 * st.m + 0.1 * g
                st.v = 0.999 * st.v + 0.001 * g * g
                tn = st.count + 1
                upd = (st.m / (1 - 0.9 ** tn)) / (np.sqrt(st.v / (1 - 0.999 ** tn)) + 1e-8) + 1e-4 * ref
                ref = ref - lr_t * upd
                st.count = tn
    assert step.item() == 6
    assert np.max(np.abs(p.cpu().numpy() - ref)) < 2e-5 * max(1.0, np.abs(ref).max())
    if kind == "adam":  # masked weights never move (weight decay, as in optax, would move them)
        mask0 = spec.flat_mask() == 0
        assert (p.cpu().numpy()[mask0] == flat[mask0]).all()


@pytest.mark.parametrize("name,N,rev,act", [("c2", 4096, True, "relu"), ("c1", 1000, True, "relu"),
                                            ("c3", 77, False, "relu"), ("ref_test", 500, True, "silu")])
def test_sample_from_noise_matches_oracle(name, N, rev, act):
    spec, flat, std, x, y = make_case(CONFIGS[name], 4, seed=6, activation=act, use_reverse=rev)
    eng = _engine(spec, std)
    rng = np.random.default_rng(7)
    u = rng.standard_normal((N, spec.n_in)).astype(np.float32)
    rows_per_obs = -(-N // 4)
    out = eng.sample_from_noise(_t(flat), _t(u), _t(y), rows_per_obs=rows_per_obs).cpu().numpy()
    ref = np.concatenate([
        onp.nde_sample_from_noise(spec, flat.astype(np.float64), std, u[i * rows_per_obs:(i + 1) * rows_per_obs].astype(np.float64), y[i])
        for i in range(4) if i * rows_per_obs < N])
    assert np.isfinite(out).all()
    assert rel_err(out, ref) < 2e-5


def test_sample_c4_parity_slice():
    """the c4 parity slice -- 65 536 supplied base-noise rows over 64 observations (1 024 each)
    through the inversion kernel against the oracle's D-pass inverse, trained-scale weights."""
    n_obs, per, = 64, 1024
    spec, flat, std, x, y = make_case(CONFIGS["c2"], n_obs, seed=16)
    eng = _engine(spec, std)
    u = np.random.default_rng(17).standard_normal((n_obs * per, spec.n_in)).astype(np.float32)
    out = eng.sample_from_noise(_t(flat), _t(u), _t(y), rows_per_obs=per).cpu().numpy()
    f64 = flat.astype(np.float64)
    ref = np.concatenate([onp.nde_sample_from_noise(spec, f64, std, u[i * per:(i + 1) * per].astype(np.float64), y[i])
                          for i in range(n_obs)])
    assert np.isfinite(out).all()
    assert rel_err(out, ref) < 2e-5


def test_sample_roundtrip_and_clamp():
    """forward(sample(u)) == u (jaxili/tests/test_nn.py:35-41 invariant), 1e-5."""
    spec, flat, std, x, y = make_case(CONFIGS["c2"], 1, seed=8, scale=0.05)
    eng = _engine(spec, std)
    N = 2048
    u = np.random.default_rng(9).standard_normal((N, spec.n_in)).astype(np.float32)
    th = eng.sample_from_noise(_t(flat), _t(u), _t(y))
    yrep = _t(np.repeat(y, N, axis=0))
    lp = eng.log_prob(_t(flat), th, yrep).cpu().numpy()
    ref = onp.nde_log_prob(spec, flat.astype(np.float64), std, th.cpu().numpy().astype(np.float64),
                           np.repeat(y, N, axis=0).astype(np.float64))
    assert rel_err(lp, ref) < TOL
    xs = (th.cpu().numpy().astype(np.float64) - std.shift) / std.scale
    ys = (np.repeat(y, N, axis=0).astype(np.float64) - std.mean) / std.std
    u_back, _ = onp.maf_forward(spec, flat.astype(np.float64), xs, ys)
    assert np.max(np.abs(u_back - u)) < 1e-4


def test_sample_philox_statistics_and_determinism():
    spec, flat, std, x, y = make_case(CONFIGS["c2"], 2, seed=10, scale=0.05)
    eng = _engine(spec, std)
    N = 1 << 16
    a = eng.sample_philox(_t(flat), N, _t(y), seed=1234)
    b = eng.sample_philox(_t(flat), N, _t(y), seed=1234)
    assert torch.equal(a, b)
    # rows are a pure function of (seed, global row): a shard starting at row_offset reproduces them
    half = eng.sample_philox(_t(flat), N // 2, _t(y[1:2]), seed=1234, row_offset=N // 2)
    assert torch.equal(half, a[N // 2:])
    assert torch.isfinite(a).all()


@pytest.mark.parametrize("name,B,act", [("c3", 1000, "relu"), ("c2", 4096, "relu"), ("c2", 2048, "silu"), ("ref_test", 40, "silu"), ("c1", 5, "tanh")])
def test_log_prob_grad_wrt_conditioner(name, B, act):
    """d log_prob / d y against torch autograd on the CPU restatement."""
    from oracle.maf_torch import TorchMaf
    spec, flat, std, x, y = make_case(CONFIGS[name], B, seed=13, activation=act)
    eng = _engine(spec, std)
    lp, dy = eng.log_prob_grad_y(_t(flat), _t(x), _t(y))
    tm = TorchMaf(spec, std, torch.float64)
    fl = torch.tensor(flat, dtype=torch.float64)
    xt = torch.tensor(x, dtype=torch.float64)
    yt = torch.tensor(y, dtype=torch.float64, requires_grad=True)
    ref = tm.log_prob(fl, xt, yt)
    ref.sum().backward()
    assert rel_err(lp.cpu().numpy(), ref.detach().numpy()) < TOL
    # per-row gradients have no batch average to hide a ReLU kink: a pre-activation within FP32 rounding of 0
    # takes the other branch than in the float64 restatement (about 1 row in 1000 at these widths), so
    # ReLU cases may have a few such rows; smooth activations must match on every row
    row_err = np.abs(dy.cpu().numpy() - yt.grad.numpy()).max(axis=1) / np.abs(yt.grad.numpy()).max()
    n_bad = int((row_err >= TOL).sum())
    assert n_bad <= (B // 250 if act == "relu" else 0), (n_bad, row_err.max())
    # the training path is unaffected by a dy call in between (work lists are per launch): bit-identical
    out = []
    for _ in range(2):
        loss, grad = torch.zeros(1, device="cuda:0"), torch.zeros(spec.n_params, device="cuda:0")
        eng.loss_grad(_t(flat), _t(x), _t(y), loss, grad)
        out.append((loss.item(), grad.cpu().numpy()))
        eng.log_prob_grad_y(_t(flat), _t(x), _t(y))
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])
    l_ref, g_ref = onp.nll_loss_and_grad(spec, flat.astype(np.float64), std, x.astype(np.float64), y.astype(np.float64))
    assert rel_err(out[1][0], l_ref) < TOL and grad_err(out[1][1], g_ref) < (1e-3 if act == "relu" else TOL)


@pytest.mark.parametrize("name,act,n_steps", [("c3", "silu", 3), ("c1", "tanh", 5), ("c3", "silu", 0)])
def test_hmc_trajectory_matches_restatement(name, act, n_steps):
    """Leapfrog steps of many chains (drift -> fused flow d/dy -> kick) against oracle/hmc_numpy.py driven by
    the float64 torch restatement of the flow's log-likelihood and its theta-gradient."""
    from oracle import hmc_numpy
    from oracle.maf_torch import TorchMaf
    n = 193
    spec, flat, std, x, y = make_case(CONFIGS[name], n, seed=5, activation=act)
    eng = _engine(spec, std)
    C = spec.n_cond
    rng = np.random.default_rng(9)
    kind = (np.arange(C) % 2).astype(np.int32)                 # alternate Normal / Uniform dimensions
    lo = np.where(kind == 1, -2.0, 0.3).astype(np.float32)
    hi = np.where(kind == 1, 2.5, 1.5).astype(np.float32)
    z0 = rng.normal(size=(n, C)).astype(np.float32)
    p0 = rng.normal(size=(n, C)).astype(np.float32)
    inv_mass = rng.uniform(0.5, 2.0, size=C).astype(np.float32)
    eps = 0.05
    x_obs = np.repeat(x[:1], n, axis=0)
    tm = TorchMaf(spec, std, torch.float64)
    fl = torch.tensor(flat, dtype=torch.float64)
    xt = torch.tensor(x_obs, dtype=torch.float64)

    def loglik_and_grad(theta):
        th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
        ll = tm.log_prob(fl, xt, th)
        ll.sum().backward()
        return ll.detach().numpy(), th.grad.numpy()

    k64, lo64, hi64 = kind.astype(np.int64), lo.astype(np.float64), hi.astype(np.float64)
    ld0, g0, th0 = hmc_numpy.log_density_and_grad(z0.astype(np.float64), k64, lo64, hi64, loglik_and_grad)
    S = {"eps": _t(np.array([eps], np.float32)), "inv_mass": _t(inv_mass),
         "prior_kind": torch.as_tensor(kind, device="cuda:0"), "lo": _t(lo), "hi": _t(hi),
         "z": _t(z0), "p": _t(p0), "g": torch.zeros(n, C, device="cuda:0"),
         "logdens": torch.zeros(n, device="cuda:0"), "theta": torch.zeros(n, C, device="cuda:0"),
         "scratch": torch.zeros(n * (C + 1), device="cuda:0")}
    eng.hmc_trajectory(_t(flat), _t(x_obs), S, 0)
    assert grad_err(S["g"].cpu().numpy(), g0) < TOL and rel_err(S["logdens"].cpu().numpy(), ld0) < TOL
    assert np.array_equal(S["z"].cpu().numpy(), z0) and np.array_equal(S["p"].cpu().numpy(), p0)
    if n_steps == 0:
        return
    z1, p1, g1, ld1, th1 = hmc_numpy.leapfrog(z0.astype(np.float64), p0.astype(np.float64), g0, eps,
                                              inv_mass.astype(np.float64), n_steps, k64, lo64, hi64, loglik_and_grad)
    eng.hmc_trajectory(_t(flat), _t(x_obs), S, n_steps, packed_current=True)
    for key, ref in (("z", z1), ("p", p1), ("g", g1), ("theta", th1)):
        assert grad_err(S[key].cpu().numpy(), ref) < 2e-5, key
    assert rel_err(S["logdens"].cpu().numpy(), ld1) < 2e-5


@pytest.mark.parametrize("kind", ["adam", "sgd", "adamw"])
@pytest.mark.parametrize("name,B", [("c2", 4096), ("c1", 128), ("c3", 3001)])
def test_train_step_equals_loss_grad_plus_clip_adam(name, B, kind):
    """mafb200_train_step (flow kernel + one cooperative tail kernel) == loss_grad followed by clip_adam,
    bit for bit: same reduction order, same update arithmetic."""
    spec, flat, std, x, y = make_case(CONFIGS[name], B, seed=21)
    hp = dict(lr=5e-3, warmup=0.1, decay_steps=50.0, clip=0.5)
    outs = []
    for fused in (False, True):
        eng = _engine(spec, std)
        opt = eng.make_opt_desc(kind=kind, weight_decay=1e-4 if kind != "adam" else 0.0, **hp)
        p = _t(flat)
        m, v, grad = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
        step = torch.zeros(1, dtype=torch.int32, device="cuda:0")
        loss = torch.zeros(1, device="cuda:0")
        xt, yt = _t(x), _t(y)
        losses = []
        for it in range(5):
            if fused:
                eng.train_step(p, xt, yt, loss, grad, m, v, step, opt, packed_current=it > 0)
            else:
                eng.loss_grad(p, xt, yt, loss, grad, packed_current=it > 0)
                eng.clip_adam(p, grad, m, v, step, opt, norm_current=True)
            losses.append(loss.item())
        outs.append((losses, p.clone(), m.clone(), v.clone(), grad.clone(), int(step.item())))
    a, b = outs
    assert a[0] == b[0] and a[5] == b[5] == 5
    for i in range(1, 5):
        assert torch.equal(a[i], b[i]), i
    assert b[0][-1] < b[0][0]


@pytest.mark.parametrize("cfg,B,act", [
    (dict(n_in=4, n_cond=3, n_layers=3, layers=[24]), 333, "relu"),             # one hidden layer: 2 masked linears
    (dict(n_in=6, n_cond=2, n_layers=2, layers=[40, 24, 16]), 129, "tanh"),     # three hidden layers: 4 masked linears
    (dict(n_in=5, n_cond=0, n_layers=3, layers=[32, 32, 32, 32]), 64, "silu"),  # unconditional, 5 masked linears
])
@pytest.mark.parametrize("flow", FLOWS)
def test_other_depths_use_the_rolled_kernel(cfg, B, act, flow):
    """`layers` of any length >= 1: the kernel is specialised for 3 masked linears per MADE and runs a
    rolled loop over the linears otherwise; forward, gradients and the conditioner gradient against the oracle."""
    spec, flat, std, x, y = make_case(cfg, B, seed=31, activation=act)
    eng = _engine(spec, std, flow=flow)
    yt = _t(y) if cfg["n_cond"] else None
    lp = eng.log_prob(_t(flat), _t(x), yt)
    ref = onp.nde_log_prob(spec, flat.astype(np.float64), std, x.astype(np.float64), y.astype(np.float64))
    assert rel_err(lp.cpu().numpy(), ref) < TOL
    loss, grad = torch.zeros(1, device="cuda:0"), torch.zeros(spec.n_params, device="cuda:0")
    eng.loss_grad(_t(flat), _t(x), yt, loss, grad)
    l_ref, g_ref = onp.nll_loss_and_grad(spec, flat.astype(np.float64), std, x.astype(np.float64), y.astype(np.float64))
    assert rel_err(loss.item(), l_ref) < TOL and grad_err(grad.cpu().numpy(), g_ref) < (1e-3 if act == "relu" else TOL)
    # forward then inverse returns the input; the log-determinants cancel (jaxili/tests/test_nn.py:35-41)
    u, ld = eng.forward(_t(flat), _t(x), yt)
    back, ld_inv = eng.inverse(_t(flat), u, yt)
    scale = np.abs(x).max()
    assert np.max(np.abs(back.cpu().numpy() - x)) < 2e-5 * max(1.0, scale) * 10
    assert torch.allclose(ld, -ld_inv, atol=1e-4)
