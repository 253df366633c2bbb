"""MCMCPosterior (jaxili/posterior/mcmc_posterior.py:30-495) with the sampler on the GPU.

The reference hands ``log_prior(theta) + log_likelihood(x, theta)`` to numpyro's NUTS/HMC, one chain per
``vmap`` lane, every leapfrog step differentiating the flow with respect to theta -- the flow's
*conditioning* input under NLE (mcmc_posterior.py:139-159, 194-201).  Here all chains are rows of one batch:
a leapfrog step is three launches (drift -> fused flow forward+reverse in d/dy mode -> kick) enqueued by
``mafb200_hmc_trajectory``; chains never leave the device.

Sampler: Hamiltonian Monte Carlo in the unconstrained space of the prior's support (as numpyro does), diagonal
mass matrix and step size adapted during warm-up (dual averaging towards ``target_accept_prob`` 0.8, windowed
variance estimates, Stan's schedule), Metropolis correction per chain.  ``"nuts_numpyro"`` and
``"hmc_numpyro"`` are accepted as method names so reference call sites keep working; both run this
fixed-length HMC (``trajectory_length`` 2*pi like numpyro's HMC, at most ``max_num_steps`` leapfrog steps,
step count jittered per transition) -- the No-U-Turn criterion is not implemented.  Parity with numpyro is
statistical only (jaxili/tests/test_posterior.py:147-166 checks shapes).

``key`` seeds torch's Philox generator (JAX's threefry stream is not reproducible here).
"""

import math
from typing import Optional

import numpy as np
import torch

from .base_posterior import NeuralPosterior

implemented_method = ["nuts_numpyro", "hmc_numpyro", "hmc_b200"]

nuts_numpyro_kwargs_default = {}
hmc_numpyro_kwargs_default = {}


class _DualAveraging:
    """Nesterov dual averaging of log(step size) (Hoffman & Gelman 2014, as in Stan/numpyro)."""

    def __init__(self, eps0, target, gamma=0.05, t0=10.0, kappa=0.75):
        self.target, self.gamma, self.t0, self.kappa = target, gamma, t0, kappa
        self.restart(eps0)

    def restart(self, eps0):
        self.mu = math.log(10.0 * eps0)
        self.h_bar, self.log_eps_bar, self.t = 0.0, 0.0, 0

    def update(self, accept_prob):
        self.t += 1
        w = 1.0 / (self.t + self.t0)
        self.h_bar = (1.0 - w) * self.h_bar + w * (self.target - accept_prob)
        log_eps = self.mu - math.sqrt(self.t) / self.gamma * self.h_bar
        eta = self.t ** (-self.kappa)
        self.log_eps_bar = eta * log_eps + (1.0 - eta) * self.log_eps_bar
        return math.exp(log_eps)

    def final(self):
        return math.exp(self.log_eps_bar)


def _warmup_windows(num_warmup):
    """(first slow iteration, ends of the slow mass-matrix windows): Stan's schedule -- init buffer 75, final
    buffer 50, base window 25 doubling; 15 % / 75 % / 10 % when the warm-up is shorter than 150."""
    if num_warmup < 20:
        return 0, []
    if num_warmup < 150:
        init, term = int(0.15 * num_warmup), int(0.1 * num_warmup)
        base = num_warmup - init - term
    else:
        init, term, base = 75, 50, 25
    ends, start, size, stop = [], init, base, num_warmup - term
    while start < stop:
        end = start + size
        if end + 2 * size > stop:
            end = stop
        ends.append(end)
        start, size = end, 2 * size
    return init, ends


class MCMCPosterior(NeuralPosterior):
    r"""Likelihood $p(x|\theta)$ learned with NLE, posterior samples by MCMC."""

    def __init__(self, model, state, prior_distr, verbose: Optional[bool] = False, x=None,
                 mcmc_method: Optional[str] = "nuts_numpyro", mcmc_kwargs: Optional[dict] = None):
        super().__init__(model, state, verbose, x)
        self.set_mcmc_method(mcmc_method)
        self.set_mcmc_kwargs(nuts_numpyro_kwargs_default if mcmc_kwargs is None else mcmc_kwargs)
        if self.verbose:
            print(f"Using MCMC method: {mcmc_method}")
            print(f"MCMC kwargs: {mcmc_kwargs}")
        self.set_prior(prior_distr)
        self.last_run = {}

    # ------------------------------------------------------------------ densities
    def log_prior(self, theta):
        """mcmc_posterior.py:120-137."""
        if self.prior_distr is None:
            raise ValueError("No prior distribution set.")
        lp = self.prior_distr.log_prob(theta)
        if len(lp.shape) > 1:
            lp = lp.sum(dim=-1)
        return lp

    @staticmethod
    def _rows(x, theta):
        x = x.reshape(1, -1) if len(x.shape) == 1 else x
        theta = theta.reshape(1, -1) if len(theta.shape) == 1 else theta
        if x.shape[0] == 1 and theta.shape[0] > 1:
            x = x.expand(theta.shape[0], -1) if isinstance(x, torch.Tensor) else np.repeat(x, theta.shape[0], axis=0)
        return x, theta

    def log_likelihood(self, x, theta):
        """mcmc_posterior.py:139-159: log p(x | theta); one x row is broadcast over the theta rows."""
        x, theta = self._rows(x, theta)
        return self.model.apply({"params": self.state.params}, x, theta, method="log_prob")

    def log_likelihood_and_grad(self, x, theta):
        """(log p(x | theta), d/d theta) from one fused forward+reverse launch (jax.grad in the reference)."""
        x, theta = self._rows(x, theta)
        return self.model.apply({"params": self.state.params}, x, theta, method="log_prob_and_grad_cond")

    def unnormalized_log_prob(self, theta, x=None):
        """mcmc_posterior.py:161-177."""
        x = self.x if x is None else x
        if x is None:
            raise ValueError("The data x must be specified or loaded in the posterior with `set_default_x()`.")
        ll = self.log_likelihood(x, theta)
        if self.prior_distr is None:      # likelihood-only use (an external sampler supplies its own prior)
            return ll
        if not isinstance(theta, torch.Tensor):
            theta = torch.as_tensor(np.asarray(theta, np.float32), device=ll.device)
        return self.log_prior(theta.to(ll.device)).to(ll.dtype) + ll

    # ------------------------------------------------------------------ sampling
    def sample(self, num_samples: int, key=None, x=None, **kwargs):
        """mcmc_posterior.py:78-118.  Returns ``(num_chains * num_samples, dim_theta)`` like
        ``mcmc.get_samples()["theta"]`` (chain-major)."""
        x = self.x if x is None else x
        if x is None:
            raise ValueError("The data x must be specified or loaded in the posterior with `set_default_x()`.")
        if self.mcmc_method not in implemented_method:
            raise NotImplementedError(
                f"The MCMC method {self.mcmc_method} is not implemented. Check print_implemented_methods().")
        if self.prior_distr is None:
            raise ValueError("A prior distribution is required to sample: set_prior(jaxili_b200.priors.Uniform(...)).")
        kw = dict(self.mcmc_kwargs)
        kw["num_samples"] = num_samples
        num_chains = int(kw.get("num_chains", 1))
        from ..model import key_to_seed
        seed = key_to_seed(key)
        initial_states = self._get_initial_state(x, num_chains, seed, **kwargs)
        return self._hmc_b200(x, seed + 1, initial_states, kw)

    def _engine(self):
        return self.model.engine(), self.model.flat_params(self.state.params)

    def _hmc_b200(self, x, seed, initial_state, kw):
        eng, flat = self._engine()
        dev = eng.device
        C, n = eng.n_cond, initial_state.shape[0]
        num_warmup = int(kw.get("num_warmup", 500))
        num_samples = int(kw.get("num_samples", 2000))
        adapt_eps = bool(kw.get("adapt_step_size", True))
        adapt_mass = bool(kw.get("adapt_mass_matrix", True))
        target = float(kw.get("target_accept_prob", 0.8))
        traj_len = float(kw.get("trajectory_length", 2.0 * math.pi))
        max_steps = int(kw.get("max_num_steps", 32))
        eps0 = float(kw.get("step_size", 1.0))

        kind, lo, hi = self.prior_distr.bounds()
        if kind.shape != (C,):
            raise ValueError(f"the prior has event shape {kind.shape}, the likelihood is conditioned on {C} parameters")
        from ..model import as_device_f32
        xr = as_device_f32(x, dev, eng.n_in)
        if xr.shape[0] != 1:
            raise ValueError("MCMC sampling conditions on a single observation x of shape (1, dim_x)")
        xr = xr.expand(n, -1).contiguous()

        f32 = dict(dtype=torch.float32, device=dev)
        S = {
            "eps": torch.full((1,), eps0, **f32), "inv_mass": torch.ones(C, **f32),
            "prior_kind": torch.as_tensor(kind, device=dev).to(torch.int32).contiguous(),
            "lo": torch.as_tensor(lo, **f32).contiguous(), "hi": torch.as_tensor(hi, **f32).contiguous(),
            "z": torch.empty(n, C, **f32), "p": torch.zeros(n, C, **f32), "g": torch.empty(n, C, **f32),
            "logdens": torch.empty(n, **f32), "theta": torch.empty(n, C, **f32),
            "scratch": torch.empty(n * (C + 1), **f32),
        }
        th0 = as_device_f32(initial_state, dev, C)
        box = S["prior_kind"] == 1
        u = ((th0 - S["lo"]) / (S["hi"] - S["lo"])).clamp(1e-6, 1.0 - 1e-6)
        S["z"].copy_(torch.where(box, torch.log(u) - torch.log1p(-u), th0))
        eng.hmc_trajectory(flat, xr, S, 0)                       # packs the weights; later calls reuse the image
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed & (2**63 - 1))
        rng = np.random.default_rng(seed & (2**63 - 1))
        keep = {k: torch.empty_like(S[k]) for k in ("z", "g", "logdens", "theta")}

        def transition(n_steps):
            """One HMC transition of every chain; returns the per-chain acceptance probabilities."""
            S["p"].normal_(generator=gen)
            S["p"].div_(S["inv_mass"].sqrt())
            for k in keep:
                keep[k].copy_(S[k])
            h0 = -S["logdens"] + 0.5 * (S["p"] ** 2 * S["inv_mass"]).sum(dim=1)
            eng.hmc_trajectory(flat, xr, S, n_steps, packed_current=True)
            h1 = -S["logdens"] + 0.5 * (S["p"] ** 2 * S["inv_mass"]).sum(dim=1)
            alpha = torch.exp((h0 - h1).clamp(max=0.0))
            alpha = torch.where(torch.isfinite(h1), alpha, torch.zeros_like(alpha))
            acc = torch.rand(n, generator=gen, device=dev) < alpha
            for k in keep:
                m = acc if S[k].dim() == 1 else acc[:, None]
                S[k].copy_(torch.where(m, S[k], keep[k]))
            return alpha

        def steps_for(eps):
            top = min(max_steps, max(1, int(math.ceil(traj_len / eps))))
            return int(rng.integers(max(1, (top + 1) // 2), top + 1))

        # ---- a reasonable first step size (Hoffman & Gelman alg. 4 on the chain-mean acceptance)
        eps = eps0
        if adapt_eps:
            direction = 0
            for _ in range(24):
                for k in keep:
                    keep[k].copy_(S[k])
                S["eps"].fill_(eps)
                a = float(transition(1).mean())
                for k in keep:          # probing must not move the chains
                    S[k].copy_(keep[k])
                d = 1 if a > target else -1
                if direction == 0:
                    direction = d
                if d != direction:
                    break
                eps = eps * 2.0 if d > 0 else eps * 0.5
        S["eps"].fill_(eps)

        # ---- warm-up: dual averaging on eps, windowed diagonal mass matrix
        da = _DualAveraging(eps, target)
        win_start, windows = _warmup_windows(num_warmup) if adapt_mass else (0, [])
        zs, zss, cnt = torch.zeros(C, dtype=torch.float64, device=dev), torch.zeros(C, dtype=torch.float64, device=dev), 0
        accept_hist = []
        for it in range(num_warmup):
            alpha = transition(steps_for(eps))
            a = float(alpha.mean())
            accept_hist.append(a)
            if adapt_eps:
                eps = min(max(da.update(a), 1e-6), 1e3)
                S["eps"].fill_(eps)
            if windows and it >= win_start:
                z64 = S["z"].double()
                zs += z64.sum(dim=0)
                zss += (z64 * z64).sum(dim=0)
                cnt += n
                if it + 1 == windows[0]:
                    windows.pop(0)
                    if cnt > 1:
                        var = (zss - zs * zs / cnt) / (cnt - 1)
                        var = (cnt / (cnt + 5.0)) * var + 1e-3 * (5.0 / (cnt + 5.0))
                        S["inv_mass"].copy_(var.float())
                    zs.zero_(), zss.zero_()
                    cnt = 0
                    if adapt_eps:
                        eps = da.final()
                        da.restart(eps)
                        S["eps"].fill_(eps)
        if adapt_eps and num_warmup > 0:
            eps = da.final()
            S["eps"].fill_(eps)

        # ---- sampling
        out = torch.empty(n, num_samples, C, **f32)
        acc_sum = torch.zeros((), **f32)
        for it in range(num_samples):
            acc_sum += transition(steps_for(eps)).mean()
            out[:, it].copy_(S["theta"])
        self.last_run = {"step_size": eps, "inv_mass": S["inv_mass"].cpu().numpy(),
                         "accept_prob": float(acc_sum) / max(num_samples, 1), "num_chains": n,
                         "warmup_accept": accept_hist}
        if self.verbose:
            print(f"HMC: step size {eps:.4g}, mean acceptance {self.last_run['accept_prob']:.3f}")
        return out.reshape(n * num_samples, C)

    # ------------------------------------------------------------------ initial states
    def _get_initial_state(self, x, num_chains: int, key, **kwargs):
        """mcmc_posterior.py:382-410: sampling-importance-resampling from the prior with posterior weights."""
        def potential_fn(theta):
            return self.unnormalized_log_prob(theta, x)

        proposal, _ = self._resample_proposal(potential_fn, key, num_draws=num_chains, **kwargs)
        return proposal

    def _resample_proposal(self, potential_fn, key, num_candidate_samples: int = 10_000, num_batches: int = 1,
                           num_draws: int = 1, **kwargs):
        """mcmc_posterior.py:412-471.  The reference draws a fresh candidate pool per chain; with thousands of
        chains that is wasteful, so one pool (never smaller than 4 candidates per chain) serves all draws."""
        from ..model import key_to_seed
        seed = key_to_seed(key)
        num_candidate_samples = max(int(num_candidate_samples), 4 * num_draws // max(num_batches, 1))
        cands, logw = [], []
        for b in range(num_batches):
            draws = self.prior_distr.sample(key=seed + 17 * b, sample_shape=(num_candidate_samples,))
            cands.append(draws)
            logw.append(potential_fn(draws))
        cands, logw = torch.cat(cands, dim=0), torch.cat(logw, dim=0)
        logw = logw - torch.logsumexp(torch.nan_to_num(logw, nan=-math.inf), dim=0)
        probs = torch.exp(logw)
        probs = torch.where(torch.isfinite(probs), probs, torch.zeros_like(probs))
        probs = probs / probs.sum()
        gen = torch.Generator(device=probs.device)
        gen.manual_seed((seed + 7919) & (2**63 - 1))
        idx = torch.multinomial(probs, num_draws, replacement=True, generator=gen)
        return cands[idx], seed + 1

    # ------------------------------------------------------------------ setters (mcmc_posterior.py:473-495)
    def set_default_x(self, x):
        self.x = x

    def set_prior(self, prior_distr):
        self.prior_distr = prior_distr

    def set_mcmc_method(self, mcmc_method: str):
        if mcmc_method not in implemented_method:
            raise NotImplementedError(
                f"The MCMC method {mcmc_method} is not implemented. Check print_implemented_methods().")
        self.mcmc_method = mcmc_method

    def set_mcmc_kwargs(self, mcmc_kwargs: dict):
        self.mcmc_kwargs = mcmc_kwargs

    def print_implemented_methods(self):
        print(f"Implemented MCMC methods: {implemented_method}")
