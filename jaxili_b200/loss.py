"""NLL losses with jaxili's signatures (jaxili/loss.py:35-83).

``loss_fn(model, params, batch) -> scalar``.  Evaluation goes through the fused forward kernel
(``model.apply(..., method="log_prob")``); ``TrainerModule`` recognises these two functions and
uses the fused forward+backward kernel for the gradient instead of autodiff (there is none here).
"""

from __future__ import annotations


def loss_nll_npe(model, params, batch):
    """-mean(log p(theta | x)); batch = (thetas, xs)  (jaxili/loss.py:35-58)."""
    thetas, xs = batch
    output = model.apply({"params": params}, thetas, xs, method="log_prob")
    return -output.mean()


def loss_nll_nle(model, params, batch):
    """-mean(log p(x | theta)); batch = (thetas, xs): arguments swapped (jaxili/loss.py:61-83)."""
    thetas, xs = batch
    return -model.apply({"params": params}, xs, thetas, method="log_prob").mean()


# name -> callable, as jaxili.inventory.func_dict.jaxili_loss_dict
jaxili_loss_dict = {"loss_nll_npe": loss_nll_npe, "loss_nll_nle": loss_nll_nle}
