"""Inference classes (jaxili.inference)."""

from .nle import NLE
from .npe import NPE, NDEDataset, default_maf_hparams
