"""NPE (jaxili/inference/npe.py): same class, methods, defaults and error behaviour; the density
estimator is the fused B200 ConditionalMAF."""

from __future__ import annotations

import warnings
from typing import Any, Callable, Dict, Iterable, Optional, Union

import numpy as np

from .. import nn as _nn
from ..loss import loss_nll_npe
from ..model import ConditionalMAF, Identity, NDE_w_Standardization, ScalarAffine, Sequential, Standardizer
from ..posterior import DirectPosterior
from ..train import TrainerModule
from ..utils import check_hparams_maf, create_data_loader, validate_theta_x

# jaxili/inference/npe.py:39-45
default_maf_hparams = {
    "n_layers": 5,
    "layers": [50, 50],
    "activation": _nn.relu,
    "use_reverse": True,
    "seed": 42,
}


class NDEDataset:
    """jaxili/inference/npe.py:48-103."""

    def __init__(self, theta, x):
        self.theta = theta
        self.x = x

    def __len__(self):
        return len(self.theta)

    def __getitem__(self, idx):
        return self.theta[idx], self.x[idx]


def split_indices(num_sims, train_test_split, key):
    """Random split as in jaxili/inference/npe.py:255-286.  ``jr.permutation`` (threefry) cannot be
    reproduced without JAX; a NumPy Generator seeded from ``key`` takes its place.  Split sizes
    follow the reference's ``int()`` arithmetic exactly."""
    is_test_set = len(train_test_split) == 3
    if is_test_set:
        train_fraction, val_fraction, test_fraction = train_test_split
        assert np.isclose(train_fraction + val_fraction + test_fraction, 1.0), \
            "The sum of the split fractions should be 1."
    elif len(train_test_split) == 2:
        train_fraction, val_fraction = train_test_split
        assert np.isclose(train_fraction + val_fraction, 1.0), "The sum of the split fractions should be 1."
    else:
        raise ValueError("train_test_split should have 2 or 3 elements.")
    if key is None:
        key = np.random.randint(0, 1000)
    from ..model import key_to_seed
    perm = np.random.default_rng(key_to_seed(key)).permutation(num_sims)
    n_tr = int(train_fraction * num_sims)
    n_tv = int((train_fraction + val_fraction) * num_sims)
    return perm[:n_tr], perm[n_tr:n_tv], (perm[n_tv:] if is_test_set else None)


class NPE:
    """Neural Posterior Estimation (jaxili/inference/npe.py:106-610)."""

    def __init__(self, model_class=ConditionalMAF, logging_level: Union[int, str] = "WARNING",
                 verbose: bool = True, model_hparams: Optional[Dict[str, Any]] = default_maf_hparams,
                 loss_fn: Callable = loss_nll_npe):
        self._model_class = model_class
        self._model_hparams = model_hparams
        self._logging_level = logging_level
        self._loss_fn = loss_fn
        self.verbose = verbose

    _nde_class = "NPE"

    def set_model_hparams(self, hparams):
        self._model_hparams = hparams

    def set_loss_fn(self, loss_fn):
        self._loss_fn = loss_fn

    def set_dataset(self, dataset, type):
        assert type in ["train", "val", "test"], "Type should be 'train', 'val' or 'test'."
        setattr(self, f"_{type}_dataset", dataset)

    def set_dataloader(self, dataloader, type):
        assert type in ["train", "val", "test"], "Type should be 'train', 'val' or 'test'."
        setattr(self, f"_{type}_dataloader", dataloader)

    def _dims(self, theta, x):
        return theta.shape[1], x.shape[1]   # NPE: density over theta, conditioned on x

    def append_simulations(self, theta, x, train_test_split: Iterable[float] = [0.7, 0.2, 0.1], key=None,
                           resident: bool = False, device=None):
        """jaxili/inference/npe.py:220-294.  ``resident=True`` (not in the reference): the simulation set is
        uploaded once, split and z-scored on the GPU and served by device-side loaders."""
        theta, x, num_sims = validate_theta_x(theta, x)
        if self.verbose:
            print("[!] Inputs are valid.")
            print(f"[!] Appending {num_sims} simulations to the dataset.")
        self._dim_params, self._dim_cond = self._dims(theta, x)
        self._num_sims = num_sims
        self._resident = bool(resident)
        if resident:
            from ..train import split_resident
            tr, va, te = split_resident(theta, x, list(train_test_split), key, device)
            self.set_dataset(tr, type="train")
            self.set_dataset(va, type="val")
            self.set_dataset(te, type="test")
            if self.verbose:
                print("[!] Dataset split into training, validation and test sets (GPU-resident).")
                print(f"[!] Training set: {len(tr)} simulations.")
                print(f"[!] Validation set: {len(va)} simulations.")
                if te is not None:
                    print(f"[!] Test set: {len(te)} simulations.")
            return self
        train_idx, val_idx, test_idx = split_indices(num_sims, list(train_test_split), key)
        self.set_dataset(NDEDataset(theta[train_idx], x[train_idx]), type="train")
        self.set_dataset(NDEDataset(theta[val_idx], x[val_idx]), type="val")
        self.set_dataset(NDEDataset(theta[test_idx], x[test_idx]) if test_idx is not None else None, type="test")
        if self.verbose:
            print("[!] Dataset split into training, validation and test sets.")
            print(f"[!] Training set: {len(train_idx)} simulations.")
            print(f"[!] Validation set: {len(val_idx)} simulations.")
            if test_idx is not None:
                print(f"[!] Test set: {len(test_idx)} simulations.")
        return self

    def _require(self, attr, msg):
        try:
            getattr(self, attr)
        except AttributeError:
            raise ValueError(msg)

    def _create_data_loader(self, **kwargs):
        """jaxili/inference/npe.py:296-340."""
        self._require("_train_dataset", "No training dataset found. Please append simulations first.")
        self._require("_val_dataset", "No validation dataset found. Please append simulations first.")
        train = [True, False] if self._test_dataset is None else [True, False, False]
        batch_size = kwargs.get("batch_size", 128)
        if self.verbose:
            print(f"[!] Creating DataLoaders with batch_size {batch_size}.")
        if getattr(self, "_resident", False):
            # same semantics as utils.create_data_loader: the train loader reshuffles every epoch (seed 42) and
            # drops the last partial batch, validation / test loaders run in order and keep it
            from ..train import ResidentLoader
            mk = lambda ds, tr: ResidentLoader(ds.theta, ds.x, batch_size, shuffle=tr, drop_last=tr,
                                               seed=kwargs.get("seed", 42), device=ds.theta.device)
            self._train_loader = mk(self._train_dataset, True)
            self._val_loader = mk(self._val_dataset, False)
            self._test_loader = None if self._test_dataset is None else mk(self._test_dataset, False)
            return
        if self._test_dataset is None:
            self._train_loader, self._val_loader = create_data_loader(
                self._train_dataset, self._val_dataset, train=train, **kwargs)
            self._test_loader = None
        else:
            self._train_loader, self._val_loader, self._test_loader = create_data_loader(
                self._train_dataset, self._val_dataset, self._test_dataset, train=train, **kwargs)

    def _stats(self, z_score_theta, z_score_x):
        """(shift, scale) of the density variable and standardizer of the conditioning variable;
        population std of the train split in FP32 (jaxili/inference/npe.py:383-399)."""
        shift = np.zeros(self._dim_params, np.float32)
        scale = np.ones(self._dim_params, np.float32)
        if getattr(self, "_resident", False):      # statistics on the device: population std, FP32
            ds = self._train_dataset
            st = lambda t: (t.mean(0).cpu().numpy(), t.std(0, unbiased=False).cpu().numpy())
            if z_score_theta:
                shift, scale = st(ds.theta)
            return shift, scale, (Standardizer(*st(ds.x)) if z_score_x else Identity())
        if z_score_theta:
            shift = np.mean(self._train_dataset.theta, axis=0, dtype=np.float32)
            scale = np.std(self._train_dataset.theta, axis=0, dtype=np.float32)
        if z_score_x:
            standardizer = Standardizer(np.mean(self._train_dataset.x, axis=0, dtype=np.float32),
                                        np.std(self._train_dataset.x, axis=0, dtype=np.float32))
        else:
            standardizer = Identity()
        return shift, scale, standardizer

    def _build_neural_network(self, z_score_theta: bool = True, z_score_x: bool = True,
                              embedding_net=None):
        """jaxili/inference/npe.py:342-413."""
        embedding_net = Identity() if embedding_net is None else embedding_net
        if self.verbose:
            print("[!] Building the neural network.")
        if self._model_class == ConditionalMAF:
            check_hparams_maf(self._model_hparams)
        else:
            warnings.warn(
                f"Model class {self._model_class} is not a base class of JaxILI.\n"
                " Check that the hyperparameters of your network are consistent.", Warning)
        self._require("_train_dataset", "No training dataset found. Please append simulations first.")
        shift, scale, standardizer = self._stats(z_score_theta, z_score_x)
        self._transformation = ScalarAffine(scale=scale, shift=shift)
        self._embedding_net = Sequential([standardizer, embedding_net])
        self._model_hparams["n_in"] = self._dim_params
        self._model_hparams["n_cond"] = self._dim_cond
        self._nde = self._model_class(**self._model_hparams)
        return NDE_w_Standardization(nde=self._nde, embedding_net=self._embedding_net,
                                     transformation=self._transformation)

    def _exmp_input(self):
        return (np.zeros((1, self._dim_params), np.float32), np.zeros((1, self._dim_cond), np.float32))

    def create_trainer(self, optimizer_hparams: Dict[str, Any], seed: int = 42,
                       logger_params: Dict[str, Any] = None, debug: bool = False,
                       check_val_every_epoch: int = 1, **kwargs):
        """jaxili/inference/npe.py:415-478."""
        try:
            self._nde
        except AttributeError:
            _ = self._build_neural_network(z_score_theta=kwargs.get("z_score_theta", True),
                                           z_score_x=kwargs.get("z_score_x", True),
                                           embedding_net=kwargs.get("embedding_net", None))
        nde_w_std_hparams = {"nde": self._nde, "embedding_net": self._embedding_net,
                             "transformation": self._transformation}
        if self.verbose:
            print("[!] Creating the Trainer module.")
        self.trainer = TrainerModule(
            model_class=NDE_w_Standardization, model_hparams=nde_w_std_hparams,
            optimizer_hparams=optimizer_hparams, loss_fn=self._loss_fn, exmp_input=self._exmp_input(),
            seed=seed, logger_params=logger_params, enable_progress_bar=self.verbose, debug=debug,
            check_val_every_epoch=check_val_every_epoch, nde_class=self._nde_class)

    def train(self, training_batch_size: int = 50, learning_rate: float = 5e-4, patience: int = 20,
              num_epochs: int = 2**31 - 1, check_val_every_epoch: int = 1, **kwargs):
        """jaxili/inference/npe.py:480-576."""
        self._require("_train_dataset", "No training dataset found. Please append simulations first.")
        try:
            self._train_loader
        except AttributeError:
            self._create_data_loader(batch_size=training_batch_size)
        if not hasattr(self, "trainer"):
            optimizer_hparams = {
                "lr": learning_rate,
                "optimizer_name": kwargs.get("optimizer_name", "adam"),
                "gradient_clip": kwargs.get("gradient_clip", 5.0),
                "warmup": kwargs.get("warmup", 0.1),
                "weight_decay": kwargs.get("weight_decay", 0.0),
            }
            logger_params = {
                "base_log_dir": kwargs.get("checkpoint_path", "checkpoints/"),
                "log_dir": kwargs.get("log_dir", None),
                "logger_type": kwargs.get("logger_type", "TensorBoard"),
            }
            self.create_trainer(optimizer_hparams=optimizer_hparams, seed=kwargs.get("seed", 42),
                                logger_params=logger_params, debug=kwargs.get("debug", False),
                                check_val_every_epoch=check_val_every_epoch, **kwargs)
            if self.verbose:
                print("[!] Training the density estimator.")
        metrics = self.trainer.train_model(self._train_loader, self._val_loader,
                                           test_loader=self._test_loader, num_epochs=num_epochs,
                                           patience=patience, min_delta=kwargs.get("min_delta", 1e-3))
        if self.verbose:
            print(f"[!] Training loss: {metrics['train/loss']}")
            print(f"[!] Validation loss: {metrics['val/loss']}")
            if self._test_loader is not None:
                print(f"[!] Test loss: {metrics['test/loss']}")
        return metrics, self.trainer.bind_model()

    def build_posterior(self, verbose: Optional[bool] = None, x=None):
        """jaxili/inference/npe.py:578-610."""
        self._require("trainer", "No trainer found. You must first create a trainer.")
        if verbose is None:
            verbose = self.verbose
        posterior = DirectPosterior(model=self.trainer.model, state=self.trainer.state, verbose=verbose, x=x)
        if self.verbose:
            print(r"[!] Posterior $p(\theta| x)$ built. The class DirectPosterior is used to sample and "
                  "evaluate the log probability.")
        return posterior
