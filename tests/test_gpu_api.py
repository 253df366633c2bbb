"""GPU tests of the reference-facing API (jaxili's own test ideas, jaxili/tests/test_nn.py,
test_npe.py, test_nle.py, test_posterior.py), plus trainer parity against the oracle."""
import json
import os

import numpy as np
import numpy.testing as npt
import pytest
import torch

from oracle import maf_numpy as onp
from tests.util import grad_err, rel_err

pytestmark = pytest.mark.gpu


def _sim(n, seed=0, D=10):
    rng = np.random.default_rng(seed)
    theta = rng.uniform(-1, 1, size=(n, D)).astype(np.float32)
    x = (theta + np.sqrt(0.1) * rng.standard_normal((n, D))).astype(np.float32)
    return theta, x


def test_conditional_maf_like_reference():
    """jaxili/tests/test_nn.py:10-61: shapes, forward->backward round trip at 1e-5, both y shapes."""
    from jaxili_b200 import nn
    from jaxili_b200.model import ConditionalMAF
    n_in, n_cond = 3, 5
    for layer, reverse in zip([3, 4], [True, False]):
        maf = ConditionalMAF(n_in, n_cond, layer, [128, 128], use_reverse=reverse, seed=42, activation=nn.silu)
        x = np.random.randn(10, n_in).astype(np.float32)
        cond = np.random.randn(10, n_cond).astype(np.float32)
        params = maf.init(0, x, cond)
        log_prob = maf.apply(params, x, cond, method="log_prob")
        assert log_prob.shape == (10,)
        u, log_det = maf.apply(params, x, cond)
        x_rec, log_det_rec = maf.apply(params, u, cond, method="backward")
        npt.assert_allclose(x, x_rec.cpu().numpy(), rtol=1e-5, atol=1e-5)
        npt.assert_allclose(log_det.cpu().numpy(), -log_det_rec.cpu().numpy(), rtol=1e-5, atol=1e-5)
        samples = maf.apply(params, cond[0], num_samples=10_000, key=0, method="sample")
        assert samples.shape == (10_000, 3)
        samples = maf.apply(params, cond[0].reshape((-1, n_cond)), num_samples=10_000, key=0, method="sample")
        assert samples.shape == (10_000, 3)
        assert torch.isfinite(samples).all()


def test_forward_matches_oracle_at_trained_scale():
    from jaxili_b200 import layout
    from jaxili_b200.model import ConditionalMAF
    maf = ConditionalMAF(4, 3, 3, [32, 16, 24], activation="tanh", use_reverse=True, seed=7)
    rng = np.random.default_rng(1)
    spec = onp.MafSpec(4, 3, 3, [32, 16, 24], activation="tanh", use_reverse=True, seed=7)
    flat = spec.init_params(rng, 0.3, np.float32)
    tree = layout.tree_from_flat(torch.tensor(flat, device="cuda:0"), maf.dims, 3)
    x = rng.standard_normal((37, 4)).astype(np.float32)
    y = rng.standard_normal((37, 3)).astype(np.float32)
    u, ld = maf.apply({"params": tree}, x, y)
    u_ref, ld_ref = onp.maf_forward(spec, flat.astype(np.float64), x.astype(np.float64), y.astype(np.float64))
    assert rel_err(u.cpu().numpy(), u_ref) < 1e-5 and rel_err(ld.cpu().numpy(), ld_ref) < 1e-5
    xr, ldr = maf.apply({"params": tree}, u_ref.astype(np.float32), y, method="backward")
    x_ref, ldr_ref = onp.maf_inverse(spec, flat.astype(np.float64), u_ref, y.astype(np.float64))
    assert rel_err(xr.cpu().numpy(), x_ref) < 2e-5 and rel_err(ldr.cpu().numpy(), ldr_ref) < 2e-5
    # a tree of plain numpy leaves (e.g. loaded from a jaxili checkpoint) is accepted too
    lp = maf.apply({"params": layout.tree_to_numpy(tree)}, x, y, method="log_prob")
    assert rel_err(lp.cpu().numpy(), onp.maf_log_prob(spec, flat.astype(np.float64), x.astype(np.float64), y.astype(np.float64))) < 1e-5


def test_network_w_standardization_like_reference():
    """jaxili/tests/test_nn.py:182-237."""
    from jaxili_b200 import nn
    from jaxili_b200.model import ConditionalMAF, Identity, NDE_w_Standardization, ScalarAffine
    n_in, n_cond = 2, 5
    rng = np.random.default_rng(0)
    theta = rng.standard_normal((10, n_in)).astype(np.float32)
    x = rng.standard_normal((10, n_cond)).astype(np.float32)
    shift = (rng.standard_normal(n_in) * 2).astype(np.float32)
    scale = rng.standard_normal(n_in).astype(np.float32)
    shifted_theta = theta * scale + shift
    maf = ConditionalMAF(n_in=n_in, n_cond=n_cond, n_layers=3, layers=[50, 50], activation=nn.relu,
                         use_reverse=True, seed=42)
    net = NDE_w_Standardization(nde=maf, embedding_net=Identity(), transformation=ScalarAffine(scale=scale, shift=shift))
    params = net.init(0, theta, x)
    npt.assert_allclose(net.apply(params, shifted_theta, method="standardize").cpu().numpy(), theta, rtol=1e-5, atol=1e-5)
    npt.assert_allclose(net.apply(params, theta, method="unstandardize").cpu().numpy(), shifted_theta, rtol=1e-5, atol=1e-5)
    npt.assert_allclose(net.apply(params, x, method="embedding").cpu().numpy(), x, rtol=1e-5, atol=1e-5)
    assert net.apply(params, theta, x, method="log_prob").shape == (10,)
    assert net.apply(params, x[0], num_samples=10_000, key=0, method="sample").shape == (10_000, n_in)
    # negative scale entries: log|scale| and the sign are handled as distrax does
    spec = onp.MafSpec(n_in, n_cond, 3, [50, 50], seed=42)
    flat = net.flat_params(params).cpu().numpy()
    std = onp.Standardization(shift, scale, np.zeros(n_cond), np.ones(n_cond))
    ref = onp.nde_log_prob(spec, flat.astype(np.float64), std, shifted_theta.astype(np.float64), x.astype(np.float64))
    assert rel_err(net.apply(params, shifted_theta, x, method="log_prob").cpu().numpy(), ref) < 1e-5


def test_trainer_steps_match_oracle(tmp_path):
    """TrainerModule.train_step (host batches) == oracle loss / clip / schedule / Adam for 5 steps."""
    from jaxili_b200 import nn
    from jaxili_b200.loss import loss_nll_npe
    from jaxili_b200.model import ConditionalMAF, Identity, NDE_w_Standardization, ScalarAffine, Sequential, Standardizer
    from jaxili_b200.train import TrainerModule
    theta, x = _sim(1024, seed=3)
    sh, sc = theta.mean(0), theta.std(0)
    me, sd = x.mean(0), x.std(0)
    maf = ConditionalMAF(10, 10, 5, [50, 50], activation=nn.relu, use_reverse=True, seed=42)
    hp = {"nde": maf, "embedding_net": Sequential([Standardizer(me, sd), Identity()]),
          "transformation": ScalarAffine(shift=sh, scale=sc)}
    tr = TrainerModule(NDE_w_Standardization, hp, {"lr": 5e-4, "gradient_clip": 5.0, "warmup": 0.1, "weight_decay": 0.0},
                       loss_nll_npe, (np.zeros((1, 10), np.float32), np.zeros((1, 10), np.float32)),
                       logger_params={"base_log_dir": str(tmp_path)}, enable_progress_bar=False)
    tr.init_optimizer(num_epochs=2**31 - 1, num_steps_per_epoch=4)
    spec = onp.MafSpec(10, 10, 5, [50, 50], seed=42)
    std = onp.Standardization(sh, sc, me, sd)
    ref = tr.state.params.flat.cpu().numpy().astype(np.float64)
    st = onp.AdamState(ref.size, np.float64)
    decay = float(int((2**31 - 1) // 2 * 4))
    for i in range(5):
        b = [theta[i * 200:(i + 1) * 200], x[i * 200:(i + 1) * 200]]
        tr.state, m = tr.train_step(tr.state, b)
        l_ref, g_ref = onp.nll_loss_and_grad(spec, ref, std, b[0].astype(np.float64), b[1].astype(np.float64))
        assert rel_err(m["loss"].item(), l_ref) < 2e-5
        ref, _, _ = onp.clip_adam_step(ref, g_ref, st, 5e-4, 0.1, decay)
    assert np.max(np.abs(tr.state.params.flat.cpu().numpy() - ref)) < 5e-6
    ev = tr.eval_step(tr.state, [theta[:300], x[:300]])
    assert rel_err(ev["loss"].item(), onp.nll_loss(spec, ref, std, theta[:300].astype(np.float64), x[:300].astype(np.float64))) < 2e-5
    # checkpoint round trip in the flax path layout
    tr.save_model(step=1)
    z = np.load(os.path.join(tr.log_dir, "checkpoint.npz"))
    assert "params/nde/layer_list_0/ConditionalMADE_0/layers_0/Dense_0/kernel" in z.files
    assert z["params/nde/layer_list_4/ConditionalMADE_0/layers_4/Dense_0/bias"].shape == (20,)
    before = tr.state.params.flat.clone()
    tr.state.params.flat.zero_()
    tr.load_model()
    assert torch.equal(tr.state.params.flat, before)


def test_graph_step_equals_eager_steps(tmp_path):
    from jaxili_b200 import nn
    from jaxili_b200.loss import loss_nll_npe
    from jaxili_b200.model import ConditionalMAF, Identity, NDE_w_Standardization, ScalarAffine, Sequential, Standardizer
    from jaxili_b200.train import TrainerModule
    theta, x = _sim(4 * 512, seed=5)

    def mk():
        maf = ConditionalMAF(10, 10, 5, [50, 50], activation=nn.relu, use_reverse=True, seed=42)
        hp = {"nde": maf, "embedding_net": Sequential([Standardizer(x.mean(0), x.std(0)), Identity()]),
              "transformation": ScalarAffine(shift=theta.mean(0), scale=theta.std(0))}
        tr = TrainerModule(NDE_w_Standardization, hp, {"lr": 1e-3, "warmup": 0.1}, loss_nll_npe,
                           (np.zeros((1, 10), np.float32), np.zeros((1, 10), np.float32)),
                           logger_params={"base_log_dir": str(tmp_path)}, enable_progress_bar=False)
        tr.init_optimizer(100, 4)
        return tr
    a, b = mk(), mk()
    step = a.make_graph_step(torch.tensor(theta, device="cuda:0"), torch.tensor(x, device="cuda:0"), 512)
    # make_graph_step ran the body twice already (warm-up + capture run nothing); count = device step
    n0 = int(a.state.step.item())
    for _ in range(6):
        step()
    n1 = int(a.state.step.item())
    assert n1 - n0 == 6
    for s in range(n1):
        i = s % 4
        b.state, _ = b.train_step(b.state, [theta[i * 512:(i + 1) * 512], x[i * 512:(i + 1) * 512]])
    assert torch.allclose(a.state.params.flat, b.state.params.flat, rtol=0, atol=1e-6)
    # several steps per graph launch: same kernels in the same order -> bit-identical to one step per launch
    c = mk()
    step4 = c.make_graph_step(torch.tensor(theta, device="cuda:0"), torch.tensor(x, device="cuda:0"), 512,
                              steps_per_launch=4)
    assert step4.steps == 4
    m0 = int(c.state.step.item())          # the warm-up run of the 4-step body
    while int(c.state.step.item()) < n1:
        step4()
    d = mk()
    step1 = d.make_graph_step(torch.tensor(theta, device="cuda:0"), torch.tensor(x, device="cuda:0"), 512)
    while int(d.state.step.item()) < int(c.state.step.item()):
        step1()
    assert m0 == 4 and int(c.state.step.item()) == int(d.state.step.item())
    assert torch.equal(c.state.params.flat, d.state.params.flat)


def test_streamed_epoch_equals_plain_steps(tmp_path):
    """train_epoch over HOST batches (double-buffered H2D on a copy stream, one CUDA graph per buffer, loss
    ring read once per epoch) == the plain train_step loop, bit for bit; numpy and pinned-tensor batches."""
    from jaxili_b200 import nn
    from jaxili_b200.loss import loss_nll_nle, loss_nll_npe
    from jaxili_b200.model import ConditionalMAF, Identity, NDE_w_Standardization, ScalarAffine, Sequential, Standardizer
    from jaxili_b200.train import TrainerModule
    theta, x = _sim(11 * 256, seed=8)

    def mk(loss, stream):
        d_in, d_c = (theta, x) if loss is loss_nll_npe else (x, theta)
        maf = ConditionalMAF(10, 10, 5, [50, 50], activation=nn.relu, use_reverse=True, seed=42)
        hp = {"nde": maf, "embedding_net": Sequential([Standardizer(d_c.mean(0), d_c.std(0)), Identity()]),
              "transformation": ScalarAffine(shift=d_in.mean(0), scale=d_in.std(0))}
        tr = TrainerModule(NDE_w_Standardization, hp, {"lr": 1e-3, "warmup": 0.1}, loss,
                           (np.zeros((1, 10), np.float32), np.zeros((1, 10), np.float32)),
                           logger_params={"base_log_dir": str(tmp_path)}, enable_progress_bar=False,
                           nde_class="NPE" if loss is loss_nll_npe else "NLE")
        tr.init_optimizer(10, 11)
        tr.stream_host_batches = stream
        return tr

    np_loader = [[theta[i * 256:(i + 1) * 256], x[i * 256:(i + 1) * 256]] for i in range(11)]
    pin_loader = [[torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()] for a, b in np_loader]
    for loss in (loss_nll_npe, loss_nll_nle):
        plain, s_np, s_pin = mk(loss, False), mk(loss, True), mk(loss, True)
        outs = []
        for tr, loader in ((plain, np_loader), (s_np, np_loader), (s_pin, pin_loader)):
            m1 = tr.train_epoch(loader)
            m2 = tr.train_epoch(loader)            # the second epoch reuses the pipeline
            outs.append((m1["train/loss"], m2["train/loss"], tr.state.params.flat.clone(), int(tr.state.step.item())))
        assert s_np._pipeline is not None and plain._pipeline is None
        for o in outs[1:]:
            assert o[3] == outs[0][3] == 22
            assert abs(o[0] - outs[0][0]) < 1e-5 and abs(o[1] - outs[0][1]) < 1e-5
            assert torch.equal(o[2], outs[0][2])
        assert outs[0][1] < outs[0][0]             # it trains
    # a ragged last batch (drop_last=False loaders) gets its own pipeline; both stay cached
    ragged = np_loader + [[theta[:100], x[:100]]]
    a, b = mk(loss_nll_npe, False), mk(loss_nll_npe, True)
    ma, mb = a.train_epoch(ragged), b.train_epoch(ragged)
    mb2, ma2 = b.train_epoch(ragged), a.train_epoch(ragged)
    assert sorted(b._pipelines) == [100, 256]
    assert abs(ma["train/loss"] - mb["train/loss"]) < 1e-5 and abs(ma2["train/loss"] - mb2["train/loss"]) < 1e-5
    assert torch.equal(a.state.params.flat, b.state.params.flat)


def test_npe_end_to_end(tmp_path):
    """jaxili/tests/test_npe.py:38-220 in one flow: split bookkeeping, loaders, standardisation,
    training smoke test (metrics, files), density estimator and DirectPosterior surface."""
    from jaxili_b200.inference import NPE, default_maf_hparams
    from jaxili_b200.model import ConditionalMAF
    from jaxili_b200.posterior import DirectPosterior
    theta, x = _sim(10_000, seed=7)
    inference = NPE(verbose=False)
    assert inference._model_class == ConditionalMAF and inference._model_hparams is default_maf_hparams
    inference = inference.append_simulations(theta, x, key=0)
    assert (inference._dim_params, inference._dim_cond, inference._num_sims) == (10, 10, 10_000)
    assert (len(inference._train_dataset), len(inference._val_dataset), len(inference._test_dataset)) == (7000, 2000, 1000)
    inference._create_data_loader(batch_size=64)
    assert next(iter(inference._train_loader))[0].shape[0] == 64
    del inference._train_loader
    model = inference._build_neural_network()
    params = model.init(0, theta, x)
    ttheta = inference._train_dataset.theta
    std_theta = (ttheta - ttheta.mean(0)) / ttheta.std(0)
    npt.assert_allclose(model.apply(params, ttheta, method="standardize").cpu().numpy(), std_theta, rtol=1e-5, atol=1e-5)
    tx = inference._train_dataset.x
    npt.assert_allclose(model.apply(params, tx, method="embedding").cpu().numpy(), (tx - tx.mean(0)) / tx.std(0), rtol=1e-5, atol=2e-5)
    assert model.apply(params, ttheta, tx, method="log_prob").shape[0] == 7000
    ckpt = str(tmp_path / "ck")
    metrics, density_estimator = inference.train(training_batch_size=256, learning_rate=2e-3, num_epochs=12,
                                                 checkpoint_path=ckpt, z_score_x=True)
    assert metrics is not None and np.isfinite(metrics["val/loss"]) and np.isfinite(metrics["train/loss"])
    assert "test/loss" in metrics
    assert density_estimator.log_prob(theta[:10], x[:10]).shape == (10,)
    assert density_estimator.sample(x[0], num_samples=10_000, key=0).shape == (10_000, 10)
    base = os.path.join(ckpt, "NDE_w_Standardization/version_0")
    assert os.path.exists(os.path.join(base, "metrics")) and os.path.exists(os.path.join(base, "hparams.json"))
    assert json.load(open(os.path.join(base, "hparams.json")))["loss_fn"] == "loss_nll_npe"
    posterior = inference.build_posterior()
    assert isinstance(posterior, DirectPosterior)
    with pytest.raises(ValueError):
        posterior.sample(10, key=0)
    posterior.set_default_x(x[0])
    s = posterior.sample(5000, key=1)
    assert s.shape == (5000, 10) and torch.isfinite(s).all()
    # training moved the loss well below the untrained flow's NLL on the same validation data
    fresh = model.apply(params, inference._val_dataset.theta, inference._val_dataset.x, method="log_prob")
    assert metrics["val/loss"] < float(-fresh.mean()) - 0.5
    lp = posterior.unnormalized_log_prob(s[:100])
    assert lp.shape == (100,)
    # chunked delivery to pinned host memory == the one-shot draw, value for value
    host = posterior.sample_to_host(5000, key=1, chunk_rows=1536)
    assert host.is_pinned() and torch.equal(host, s.cpu())
    with pytest.raises(ValueError):
        posterior.unnormalized_log_prob(s[:100], x=x[:7])


def test_nle_end_to_end(tmp_path):
    from jaxili_b200.inference import NLE
    rng = np.random.default_rng(2)
    theta = rng.uniform(-3, 3, size=(4000, 5)).astype(np.float32)
    x = (np.repeat(theta[:, :2], 4, axis=1) + 0.3 * rng.standard_normal((4000, 8))).astype(np.float32)
    inf = NLE(verbose=False).append_simulations(theta, x, key=1)
    assert (inf._dim_params, inf._dim_cond) == (8, 5)
    metrics, de = inf.train(training_batch_size=200, num_epochs=2, checkpoint_path=str(tmp_path / "nle"))
    assert np.isfinite(metrics["val/loss"])
    like = inf.build_posterior(x=x[0])
    ll = like.log_likelihood(x[0], theta[:50])
    assert ll.shape == (50,) and torch.isfinite(ll).all()
    # training lowers the NLL: true theta explains x[0] better than a far-away theta
    far = theta[:1].copy() + 2.5
    assert like.log_likelihood(x[0], theta[:1]).item() > like.log_likelihood(x[0], far).item()
    with pytest.raises(ValueError):
        like.sample(10)                     # no prior set
    # MCMC over the learned likelihood (jaxili/tests/test_posterior.py:102-166): shapes, support, and
    # agreement of the chains with a self-normalised importance-sampling estimate from the same density
    from jaxili_b200.priors import Uniform
    prior = Uniform(-3.0 * np.ones(5, np.float32), 3.0 * np.ones(5, np.float32))
    post = inf.build_posterior(prior_distr=prior, x=x[0].reshape(1, -1))
    assert post.mcmc_method == "nuts_numpyro" and post.mcmc_kwargs == {}
    assert post.log_prior(torch.as_tensor(theta[:10])).shape == (10,)
    assert post.unnormalized_log_prob(theta[:10], x[:10]).shape == (10,)
    post.set_mcmc_kwargs({"num_warmup": 60})
    s1 = post.sample(num_samples=100, key=0)
    assert s1.shape == (100, 5) and torch.isfinite(s1).all()
    post.set_mcmc_method("hmc_numpyro")
    post.set_mcmc_kwargs({"num_chains": 512, "num_warmup": 200})
    s = post.sample(num_samples=40, key=3)
    assert s.shape == (512 * 40, 5) and torch.isfinite(s).all()
    assert (s > -3).all() and (s < 3).all()
    assert 0.5 < post.last_run["accept_prob"] < 0.99
    cand = prior.sample(key=11, sample_shape=(1 << 20,))
    w = torch.softmax(post.unnormalized_log_prob(cand).double(), dim=0)
    is_mean = (w[:, None] * cand.double()).sum(0)
    is_std = ((w[:, None] * (cand.double() - is_mean) ** 2).sum(0)).sqrt()
    ess = 1.0 / float((w * w).sum())
    assert ess > 500, ess
    assert torch.allclose(s.double().mean(0), is_mean, atol=0.12), (s.mean(0), is_mean)
    assert torch.allclose(s.double().std(0), is_std, rtol=0.15, atol=0.03), (s.std(0), is_std)
    with pytest.raises(NotImplementedError):
        post.set_mcmc_method("wrong_method")


def test_errors_like_reference():
    from jaxili_b200.inference import NPE
    from jaxili_b200.model import ConditionalMAF
    inf = NPE(verbose=False)
    with pytest.raises(ValueError):
        inf.train()
    with pytest.raises(ValueError):
        inf.build_posterior()
    with pytest.raises(AssertionError):
        inf.append_simulations(np.zeros((5, 2)), np.zeros((5, 2), np.float32))   # float64 theta
    with pytest.raises(AssertionError):
        inf.append_simulations(np.zeros((5, 2), np.float32), np.zeros((6, 2), np.float32))
    with pytest.raises(ValueError):
        ConditionalMAF(2, 2, 2, [8], activation="not_an_activation")
    with pytest.raises(ValueError):
        ConditionalMAF(1, 2, 2, [8], activation="relu", seed=1).engine()   # n_in=1: randint(0,0) in jaxili


def test_resident_loader_and_on_device_split(tmp_path):
    """GPU-resident simulation set, split / z-scored / shuffled on the device.  Split sizes follow the
    reference's int() arithmetic (jaxili/inference/npe.py:269-280), the loaders keep create_data_loader's
    semantics (jaxili/utils.py:72-81), and training runs on batches that are views of the resident tensors --
    including a width (3) and batch size (50) whose row slices are not 16-byte aligned."""
    from jaxili_b200.inference import NPE
    from jaxili_b200.train import ResidentLoader
    rng = np.random.default_rng(3)
    n = 4001
    theta = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    x = (theta @ rng.standard_normal((3, 5)).astype(np.float32) + 0.2 * rng.standard_normal((n, 5))).astype(np.float32)
    inf = NPE(verbose=False, model_hparams=dict(n_layers=3, layers=[32, 32], activation="relu", use_reverse=True, seed=42))
    inf.append_simulations(theta, x, train_test_split=[0.7, 0.2, 0.1], key=5, resident=True)
    tr, va, te = inf._train_dataset, inf._val_dataset, inf._test_dataset
    assert (len(tr), len(va), len(te)) == (int(0.7 * n), int(0.9 * n) - int(0.7 * n), n - int(0.9 * n))
    assert tr.theta.is_cuda and tr.x.is_cuda
    # the three parts are a partition of the simulation set (rows kept together)
    rows = torch.cat([torch.cat([d.theta, d.x], 1) for d in (tr, va, te)]).cpu().numpy()
    full = np.concatenate([theta, x], 1)
    assert rows.shape == full.shape
    assert np.array_equal(rows[np.lexsort(rows.T[::-1])], full[np.lexsort(full.T[::-1])])
    # same key -> same split; statistics = population mean / std of the train part
    inf2 = NPE(verbose=False).append_simulations(theta, x, key=5, resident=True)
    assert torch.equal(inf2._train_dataset.theta, tr.theta)
    shift, scale, stdz = inf._stats(True, True)
    npt.assert_allclose(shift, tr.theta.cpu().numpy().mean(0), rtol=1e-5, atol=1e-6)
    npt.assert_allclose(scale, tr.theta.cpu().numpy().std(0), rtol=1e-5, atol=1e-6)
    # loader semantics
    inf._create_data_loader(batch_size=50)
    tl, vl = inf._train_loader, inf._val_loader
    assert isinstance(tl, ResidentLoader) and len(tl) == len(tr) // 50 and len(vl) == -(-len(va) // 50)
    e1 = torch.cat([b[0] for b in tl]); e2 = torch.cat([b[0] for b in tl])
    assert e1.shape[0] == len(tl) * 50 and not torch.equal(e1, e2)              # reshuffled every epoch
    key = lambda t: t[:, 0].sort().values
    sub = torch.isin(key(e1), key(tr.theta)).all()                              # every batch row is a train row
    assert bool(sub)
    tl_b = ResidentLoader(tr.theta, tr.x, 50, shuffle=True, drop_last=True, seed=42)
    assert torch.equal(torch.cat([b[0] for b in tl_b]), e1)                     # seeded: same first epoch
    assert torch.equal(torch.cat([b[0] for b in vl]), va.theta)                 # validation: in order, last batch kept
    # training on the resident loaders (device batches, misaligned slices: 50 rows x 3 floats = 600 B)
    metrics, est = inf.train(training_batch_size=50, learning_rate=2e-3, num_epochs=6, checkpoint_path=str(tmp_path / "ck"))
    assert np.isfinite(metrics["train/loss"]) and np.isfinite(metrics["val/loss"]) and "test/loss" in metrics
    assert est.log_prob(theta[:7], x[:7]).shape == (7,)


def test_eval_model_on_device_equals_host_loop():
    """eval_model's size-weighted mean (jaxili/train.py:453-487) accumulated on the device equals the
    host loop over eval_step, ragged last batch included."""
    from jaxili_b200.inference import NPE
    theta, x = _sim(3000, seed=9)
    inf = NPE(verbose=False).append_simulations(theta, x, key=1)
    inf._create_data_loader(batch_size=128)
    inf.create_trainer({"lr": 1e-3, "optimizer_name": "adam", "gradient_clip": 5.0, "warmup": 0.1, "weight_decay": 0.0},
                       logger_params={"base_log_dir": "/tmp/jaxili_b200_eval_test"})
    tr = inf.trainer
    tr.init_optimizer(3, len(inf._train_loader))
    tr.train_epoch(inf._train_loader)
    dev = tr.eval_model(inf._val_loader, log_prefix="val/")["val/loss"]
    tot, cnt = 0.0, 0
    for batch in inf._val_loader:
        m = tr.eval_step(tr.state, batch)
        tot += float(m["loss"].item()) * batch[0].shape[0]
        cnt += batch[0].shape[0]
    assert abs(dev - tot / cnt) < 1e-5 * max(1.0, abs(dev))
    # a call on OTHER parameters re-packs the shared handle; the next training step must not run on them
    eng = tr.model.engine()
    other = tr.state.params.flat.clone().mul_(0.5)
    xb, yb = tr._batch_xy(next(iter(inf._val_loader)))
    eng.log_prob(other, torch.tensor(xb, device=eng.device), torch.tensor(yb, device=eng.device))
    before = tr.state.params.flat.clone()
    ref_loss = tr.eval_model(inf._val_loader)["loss"]
    assert abs(ref_loss - dev) < 1e-5 * max(1.0, abs(dev))                      # eval packs the trainer's weights again
    eng.log_prob(other, torch.tensor(xb, device=eng.device), torch.tensor(yb, device=eng.device))
    tr.stream_host_batches = False
    tr.state, m = tr.train_step(tr.state, next(iter(inf._train_loader)))        # packed_current=True would be stale here
    assert abs(float(m["loss"].item()) - ref_loss) < 0.5                        # loss of the trainer's weights, not of `other`
    assert not torch.equal(before, tr.state.params.flat)
