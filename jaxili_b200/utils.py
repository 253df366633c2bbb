"""Validators and data-loading helpers with jaxili's names (jaxili/utils.py).

``create_data_loader`` keeps the reference semantics that affect results: the train
loader reshuffles every epoch from ``torch.Generator().manual_seed(seed)`` and drops the last
partial batch; validation / test loaders keep order and the last partial batch.  It is an
in-process loader over numpy arrays (no worker processes: the 4-process torch DataLoader is the
reference's real bottleneck); batches are lists ``[theta, x]`` of float32 numpy
arrays exactly like ``numpy_collate`` produces.  ``ResidentLoader`` (train.py) is the GPU-resident
variant used at kernel speed.
"""

from __future__ import annotations

from typing import Any, Sequence, Union

import numpy as np
import torch


def numpy_collate(batch):
    """jaxili/utils.py:17-41."""
    if isinstance(batch[0], np.ndarray):
        return np.stack(batch)
    elif isinstance(batch[0], (tuple, list)):
        transposed = zip(*batch)
        return [numpy_collate(samples) for samples in transposed]
    else:
        return np.array(batch)


class NumpyLoader:
    """Minimal DataLoader over a dataset exposing ``.theta`` and ``.x`` (or generic indexing)."""

    def __init__(self, dataset, batch_size=128, shuffle=False, drop_last=False, seed=42):
        self.dataset, self.batch_size = dataset, int(batch_size)
        self.shuffle, self.drop_last = shuffle, drop_last
        self.generator = torch.Generator().manual_seed(seed)

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator).numpy() if self.shuffle else np.arange(n)
        for i in range(len(self)):
            idx = order[i * self.batch_size:(i + 1) * self.batch_size]
            if hasattr(self.dataset, "theta") and hasattr(self.dataset, "x"):
                yield [np.asarray(self.dataset.theta)[idx], np.asarray(self.dataset.x)[idx]]
            else:
                yield numpy_collate([self.dataset[int(j)] for j in idx])


def create_data_loader(*datasets, train: Union[bool, Sequence[bool]] = True, batch_size: int = 128,
                       num_workers: int = 4, seed: int = 42):
    """jaxili/utils.py:44-83 (``num_workers`` accepted and ignored: in-process loader)."""
    loaders = []
    if not isinstance(train, (list, tuple)):
        train = [train for _ in datasets]
    for dataset, is_train in zip(datasets, train):
        loaders.append(NumpyLoader(dataset, batch_size=batch_size, shuffle=is_train, drop_last=is_train,
                                   seed=seed))
    return loaders


def check_density_estimator(estimator_arg: str):
    """jaxili/utils.py:86-98."""
    if estimator_arg not in ["maf", "mdn", "realnvp"]:
        raise ValueError(
            f"Invalid density estimator argument: {estimator_arg}. Options are 'maf', 'mdn' and 'realnvp'.")


def validate_theta_x(theta: Any, x: Any):
    """jaxili/utils.py:101-136: arrays, same batch size, float32."""
    ok = (np.ndarray, torch.Tensor)
    assert isinstance(theta, ok), "theta should be a jax array."
    assert isinstance(x, ok), "x should be a jax array."
    assert theta.shape[0] == x.shape[0], (
        f"Number of parameter sets ({theta.shape[0]}) and number of simulation outputs "
        f"({x.shape[0]}) should be the same.")
    assert theta.dtype in (np.float32, torch.float32), "theta should have dtype float32."
    assert x.dtype in (np.float32, torch.float32), "x should have dtype float32."
    if isinstance(theta, torch.Tensor):
        theta = theta.detach().cpu().numpy()
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return theta, x, theta.shape[0]


def check_hparams_maf(hparams: dict):
    """jaxili/utils.py:139-156."""
    assert "n_layers" in hparams, "n_layers not found in hyperparameters."
    assert "layers" in hparams, "layers not found in hyperparameters."
    assert "activation" in hparams, "activation not found in hyperparameters."
    assert "use_reverse" in hparams, "use_reverse not found in hyperparameters."
    assert isinstance(hparams["n_layers"], int), "n_layers should be an int."
    assert isinstance(hparams["layers"], list), "layers should be a list of int."
    assert callable(hparams["activation"]) or isinstance(hparams["activation"], str), \
        "activation should be a callable."
    assert isinstance(hparams["use_reverse"], bool), "use_reverse should be a boolean."


def handle_non_serializable(obj):
    """jaxili/utils.py:195-206."""
    return str(obj)
