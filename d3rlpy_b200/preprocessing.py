"""Scalers on the update path (d3rlpy/preprocessing/scalers.py): StandardScaler and PixelScaler.
`transform` runs fused inside the gather / staging kernels; these classes only hold parameters."""
from __future__ import annotations

import numpy as np


class StandardScaler:
    """(x - mean) / (std + eps), eps=1e-3 (scalers.py:256-354)."""

    TYPE = "standard"

    def __init__(self, dataset=None, mean=None, std=None, eps: float = 1e-3):
        self._mean = None if mean is None else np.asarray(mean, dtype=np.float64)
        self._std = None if std is None else np.asarray(std, dtype=np.float64)
        self._eps = eps
        if dataset is not None:
            self.fit_dataset(dataset)

    def fit_dataset(self, dataset) -> None:
        """Statistics over every transition's observation (scalers.py:318-343): float64 mean and
        population std over transitions (the dropped last step of truncated episodes is excluded)."""
        if self._mean is not None and self._std is not None:
            return
        obs = dataset.transition_observations().astype(np.float64)
        self._mean = obs.mean(axis=0)
        self._std = np.sqrt(((obs - self._mean) ** 2).mean(axis=0))


class PixelScaler:
    """x / 255 (scalers.py:66-110); fused into the first conv layer's load."""

    TYPE = "pixel"
