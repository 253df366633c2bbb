"""NLE (jaxili/inference/nle.py): the flow models p(x | theta); roles of theta and x are swapped
relative to NPE (dims 188-189, standardisation 323-341, ``nde_class="NLE"`` 421)."""

from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Union

import numpy as np

from ..loss import loss_nll_nle
from ..model import ConditionalMAF, Identity, Standardizer
from .npe import NPE, default_maf_hparams


class NLE(NPE):
    """Neural Likelihood Estimation with jaxili's surface."""

    _nde_class = "NLE"

    def __init__(self, model_class=ConditionalMAF, logging_level: Union[int, str] = "WARNING",
                 verbose: bool = True, model_hparams: Optional[Dict[str, Any]] = default_maf_hparams,
                 loss_fn: Callable = loss_nll_nle):
        super().__init__(model_class, logging_level, verbose, model_hparams, loss_fn)

    def _dims(self, theta, x):
        return x.shape[1], theta.shape[1]   # the distribution of the data is learned, conditioned on theta

    def _stats(self, z_score_theta, z_score_x):
        """jaxili/inference/nle.py:323-341: Standardizer on theta (conditioner), ScalarAffine on x."""
        if getattr(self, "_resident", False):      # statistics on the device: population std, FP32
            ds = self._train_dataset
            st = lambda t: (t.mean(0).cpu().numpy(), t.std(0, unbiased=False).cpu().numpy())
            shift, scale = st(ds.x) if z_score_x else (np.zeros(self._dim_params, np.float32),
                                                       np.ones(self._dim_params, np.float32))
            return shift, scale, (Standardizer(*st(ds.theta)) if z_score_theta else Identity())
        if z_score_theta:
            standardizer = Standardizer(np.mean(self._train_dataset.theta, axis=0, dtype=np.float32),
                                        np.std(self._train_dataset.theta, axis=0, dtype=np.float32))
        else:
            standardizer = Identity()
        shift = np.zeros(self._dim_params, np.float32)
        scale = np.ones(self._dim_params, np.float32)
        if z_score_x:
            shift = np.mean(self._train_dataset.x, axis=0, dtype=np.float32)
            scale = np.std(self._train_dataset.x, axis=0, dtype=np.float32)
        return shift, scale, standardizer

    def _exmp_input(self):
        # jaxili/inference/nle.py:403-406: (cond, params); TrainerModule reverses it for NLE
        return (np.zeros((1, self._dim_cond), np.float32), np.zeros((1, self._dim_params), np.float32))

    def build_posterior(self, prior_distr=None, verbose: Optional[bool] = None, x=None,
                        mcmc_method: Optional[str] = "nuts_numpyro", mcmc_kwargs: Optional[Dict[str, Any]] = None):
        """jaxili/inference/nle.py:522-574: wrap the trained likelihood in an ``MCMCPosterior``.  The sampler
        is the vectorised-chains HMC of ``jaxili_b200.posterior.mcmc_posterior`` (numpyro is not on this
        path); ``prior_distr`` is a ``jaxili_b200.priors`` distribution (``Uniform`` / ``Normal``), needed
        only once ``sample`` is called."""
        self._require("trainer", "No trainer found. You must first create a trainer.")
        from ..posterior.mcmc_posterior import MCMCPosterior
        return MCMCPosterior(model=self.trainer.model, state=self.trainer.state, prior_distr=prior_distr,
                             verbose=self.verbose if verbose is None else verbose, x=x,
                             mcmc_method=mcmc_method, mcmc_kwargs=mcmc_kwargs)
