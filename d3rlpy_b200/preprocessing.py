"""Scalers on the update path (d3rlpy/preprocessing/{scalers,action_scalers,reward_scalers}.py).

`transform` never runs on the host: observation scalers are fused into the gather kernels (device-sampled batches) or
run as `d3b_standardize` on a host-staged batch; action and reward scalers run as `d3b_scale_actions` /
`d3b_scale_rewards` on the minibatch buffer; `reverse_transform` of predicted actions is `d3b_unscale_actions`
(d3rlpy/torch_utility.py:179-185, d3rlpy/algos/torch/base.py:52-80).  These classes hold the parameters, restate the
reference's `fit`, and describe the transform as the few float32 constants the kernels take.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np


def _resolve(data):
    """(MDPDataset, transition indices) of what `fit` was given: an MDPDataset, a list of Transitions / Episodes of
    one MDPDataset (base.py:494-507), or the `TransitionSubset` that `AlgoBase.fit` builds from such a list."""
    if hasattr(data, "_meta"):
        return data, np.arange(data._meta.shape[0])
    if hasattr(data, "_t_index"):
        return data._ds, data._t_index
    ds = data[0]._ds
    assert all(x._ds is ds for x in data), "transitions must come from one MDPDataset"
    if hasattr(data[0], "_tr"):   # Episodes: all their transitions, in order
        return ds, np.concatenate([np.arange(*e._tr) for e in data]).astype(np.int64)
    return ds, np.fromiter((tr._t for tr in data), dtype=np.int64, count=len(data))


class TransitionSubset:
    """The transitions a `fit` call trains on, as indices into one MDPDataset's transition table."""

    def __init__(self, data):
        self._ds, self._t_index = _resolve(data)

    def __len__(self):
        return len(self._t_index)


def _transition_arrays(data):
    """(observations, actions, rewards, episode id) of every transition passed."""
    ds, t = _resolve(data)
    step = ds._meta[t, 0]
    return ds._observations[step], ds._actions[step], ds._rewards[step], ds._meta[t, 1]


class _Scaler:
    TYPE = "none"

    def get_type(self) -> str:
        return self.TYPE

    def fit_dataset(self, dataset) -> None:
        self.fit(dataset)

    def fit(self, transitions) -> None:
        pass

    def fit_with_env(self, env) -> None:
        pass

    def get_params(self, deep: bool = False) -> Dict[str, Any]:
        return {}


# --------------------------------------------------------------------------------------- observation scalers
class StandardScaler(_Scaler):
    """(x - mean) / (std + eps), eps=1e-3 (scalers.py:256-366)."""

    TYPE = "standard"

    def __init__(self, dataset=None, mean=None, std=None, eps: float = 1e-3):
        self._mean = None if mean is None else np.asarray(mean, dtype=np.float64)
        self._std = None if std is None else np.asarray(std, dtype=np.float64)
        self._eps = eps
        if dataset is not None:
            self.fit_dataset(dataset)

    def fit(self, transitions) -> None:
        """Statistics over every transition's observation (scalers.py:318-343): float64 mean and
        population std over transitions (the dropped last step of truncated episodes is excluded)."""
        if self._mean is not None and self._std is not None:
            return
        obs = _transition_arrays(transitions)[0].astype(np.float64)
        self._mean = obs.mean(axis=0)
        self._std = np.sqrt(((obs - self._mean) ** 2).mean(axis=0))

    def fit_with_env(self, env) -> None:
        if self._mean is not None and self._std is not None:
            return
        raise NotImplementedError("standard scaler does not support fit_with_env.")   # scalers.py:345-350

    def affine_f32(self):
        """(subtrahend, divisor, eps) of `(x - s) / (d + eps)` as the kernels take them."""
        assert self._mean is not None and self._std is not None, "standard scaler is not fitted"   # scalers.py:351
        return (np.asarray(self._mean, np.float32).reshape(-1), np.asarray(self._std, np.float32).reshape(-1),
                float(self._eps))

    def get_params(self, deep: bool = False) -> Dict[str, Any]:
        return {"mean": self._mean, "std": self._std, "eps": self._eps}


class MinMaxScaler(_Scaler):
    """(x - min) / (max - min) (scalers.py:119-254); the divisor is the float32 difference of the float32 bounds, as
    `transform` computes it on tensors (:209-218)."""

    TYPE = "min_max"

    def __init__(self, dataset=None, maximum=None, minimum=None):
        self._minimum = None if minimum is None else np.asarray(minimum)
        self._maximum = None if maximum is None else np.asarray(maximum)
        if maximum is None or minimum is None:
            self._minimum = self._maximum = None
        if dataset is not None:
            self.fit_dataset(dataset)

    def fit(self, transitions) -> None:
        if self._minimum is not None and self._maximum is not None:
            return
        obs = _transition_arrays(transitions)[0]
        self._minimum = obs.min(axis=0).reshape((1,) + obs.shape[1:])
        self._maximum = obs.max(axis=0).reshape((1,) + obs.shape[1:])

    def fit_with_env(self, env) -> None:
        if self._minimum is not None and self._maximum is not None:
            return
        shape = env.observation_space.shape
        self._minimum = np.asarray(env.observation_space.low).reshape((1,) + shape)
        self._maximum = np.asarray(env.observation_space.high).reshape((1,) + shape)

    def affine_f32(self):
        assert self._minimum is not None and self._maximum is not None, "min_max scaler is not fitted"   # scalers.py:210
        mn = np.asarray(self._minimum, np.float32).reshape(-1)
        mx = np.asarray(self._maximum, np.float32).reshape(-1)
        return mn, mx - mn, 0.0

    def get_params(self, deep: bool = False) -> Dict[str, Any]:
        return {"maximum": self._maximum, "minimum": self._minimum}


class PixelScaler(_Scaler):
    """x / 255 (scalers.py:66-110); fused into the first conv layer's load."""

    TYPE = "pixel"


# --------------------------------------------------------------------------------------- action scaler
class MinMaxActionScaler(_Scaler):
    """((a - min) / (max - min)) * 2 - 1 (action_scalers.py:139-212)."""

    TYPE = "min_max"

    def __init__(self, dataset=None, maximum=None, minimum=None):
        self._minimum = None if minimum is None else np.asarray(minimum)
        self._maximum = None if maximum is None else np.asarray(maximum)
        if maximum is None or minimum is None:
            self._minimum = self._maximum = None
        if dataset is not None:
            self.fit_dataset(dataset)

    def fit(self, transitions) -> None:
        if self._minimum is not None and self._maximum is not None:
            return
        act = _transition_arrays(transitions)[1]
        self._minimum = act.min(axis=0).reshape((1,) + act.shape[1:])
        self._maximum = act.max(axis=0).reshape((1,) + act.shape[1:])

    def fit_with_env(self, env) -> None:
        if self._minimum is not None and self._maximum is not None:
            return
        shape = env.action_space.shape
        self._minimum = np.asarray(env.action_space.low).reshape((1,) + shape)
        self._maximum = np.asarray(env.action_space.high).reshape((1,) + shape)

    def bounds_f32(self):
        assert self._minimum is not None and self._maximum is not None, "action scaler is not fitted"   # action_scalers.py:186
        return (np.asarray(self._minimum, np.float32).reshape(-1), np.asarray(self._maximum, np.float32).reshape(-1))

    def get_params(self, deep: bool = False) -> Dict[str, Any]:
        return {"minimum": self._minimum, "maximum": self._maximum}


# --------------------------------------------------------------------------------------- reward scalers
class _RewardScaler(_Scaler):
    """Every reward scaler of the reference is `(mul * (clamp(r, lo, hi) - sub)) / div` for some constants."""

    def constants(self):
        """(lo, hi, sub, mul, div) as python floats (rounded to float32 by the C call, like torch rounds the python
        scalars it combines with a float32 tensor)."""
        raise NotImplementedError


class MultiplyRewardScaler(_RewardScaler):
    """multiplier * r (reward_scalers.py:96-135)."""

    TYPE = "multiply"

    def __init__(self, multiplier: Optional[float] = None):
        self._multiplier = multiplier

    def constants(self):
        return (-np.inf, np.inf, 0.0, float(self._multiplier), 1.0)

    def get_params(self, deep: bool = False) -> Dict[str, Any]:
        return {"multiplier": self._multiplier}


class ClipRewardScaler(_RewardScaler):
    """multiplier * clamp(r, low, high) (reward_scalers.py:138-190)."""

    TYPE = "clip"

    def __init__(self, low: Optional[float] = None, high: Optional[float] = None, multiplier: float = 1.0):
        self._low, self._high, self._multiplier = low, high, multiplier

    def constants(self):
        lo = -np.inf if self._low is None else float(self._low)
        hi = np.inf if self._high is None else float(self._high)
        return (lo, hi, 0.0, float(self._multiplier), 1.0)

    def get_params(self, deep: bool = False) -> Dict[str, Any]:
        return {"low": self._low, "high": self._high, "multiplier": self._multiplier}


class MinMaxRewardScaler(_RewardScaler):
    """multiplier * (r - min) / (max - min) (reward_scalers.py:193-281)."""

    TYPE = "min_max"

    def __init__(self, dataset=None, minimum: Optional[float] = None, maximum: Optional[float] = None,
                 multiplier: float = 1.0):
        self._minimum = self._maximum = None
        self._multiplier = multiplier
        if dataset is not None:
            self.fit_dataset(dataset)
        elif minimum is not None and maximum is not None:
            self._minimum, self._maximum = minimum, maximum

    def fit(self, transitions) -> None:
        if self._minimum is not None and self._maximum is not None:
            return
        rewards = _transition_arrays(transitions)[2]
        self._minimum, self._maximum = float(np.min(rewards)), float(np.max(rewards))

    def constants(self):
        assert self._minimum is not None and self._maximum is not None, "reward scaler is not fitted"   # reward_scalers.py:262
        return (-np.inf, np.inf, float(self._minimum), float(self._multiplier), float(self._maximum - self._minimum))

    def get_params(self, deep: bool = False) -> Dict[str, Any]:
        return {"minimum": self._minimum, "maximum": self._maximum, "multiplier": self._multiplier}


class StandardRewardScaler(_RewardScaler):
    """multiplier * (r - mean) / (std + eps) (reward_scalers.py:284-376)."""

    TYPE = "standard"

    def __init__(self, dataset=None, mean: Optional[float] = None, std: Optional[float] = None, eps: float = 1e-3,
                 multiplier: float = 1.0):
        self._mean = self._std = None
        self._eps, self._multiplier = eps, multiplier
        if dataset is not None:
            self.fit_dataset(dataset)
        elif mean is not None and std is not None:
            self._mean, self._std = mean, std

    def fit(self, transitions) -> None:
        if self._mean is not None and self._std is not None:
            return
        rewards = _transition_arrays(transitions)[2].astype(np.float64)  # list of python floats in the reference
        self._mean, self._std = float(np.mean(rewards)), float(np.std(rewards))

    def constants(self):
        assert self._mean is not None and self._std is not None, "reward scaler is not fitted"   # reward_scalers.py:357
        return (-np.inf, np.inf, float(self._mean), float(self._multiplier), float(self._std + self._eps))

    def get_params(self, deep: bool = False) -> Dict[str, Any]:
        return {"mean": self._mean, "std": self._std, "eps": self._eps, "multiplier": self._multiplier}


class ReturnBasedRewardScaler(_RewardScaler):
    """multiplier * r / (return_max - return_min) over the episode returns of the data (reward_scalers.py:379-488)."""

    TYPE = "return"

    def __init__(self, dataset=None, return_max: Optional[float] = None, return_min: Optional[float] = None,
                 multiplier: float = 1.0):
        self._return_max = self._return_min = None
        self._multiplier = multiplier
        if dataset is not None:
            self.fit_dataset(dataset)
        elif return_max is not None and return_min is not None:
            self._return_max, self._return_min = return_max, return_min

    def fit(self, transitions) -> None:
        """Whole-episode sums of transition rewards, whichever transitions of an episode were passed (:441-463)."""
        if self._return_max is not None and self._return_min is not None:
            return
        ds = _resolve(transitions)[0]
        episodes = np.unique(_transition_arrays(transitions)[3])
        start = ds._meta[:, 1]
        returns = []
        for s in episodes:
            t = np.nonzero(start == s)[0]
            returns.append(float(np.sum(ds._rewards[ds._meta[t, 0]].astype(np.float64))))
        self._return_max, self._return_min = float(np.max(returns)), float(np.min(returns))

    def constants(self):
        assert self._return_max is not None and self._return_min is not None, "reward scaler is not fitted"   # :471
        return (-np.inf, np.inf, 0.0, float(self._multiplier), float(self._return_max - self._return_min))

    def get_params(self, deep: bool = False) -> Dict[str, Any]:
        return {"return_max": self._return_max, "return_min": self._return_min, "multiplier": self._multiplier}


# --------------------------------------------------------------------------------------- registries
SCALER_LIST = {c.TYPE: c for c in (PixelScaler, MinMaxScaler, StandardScaler)}
ACTION_SCALER_LIST = {c.TYPE: c for c in (MinMaxActionScaler,)}
REWARD_SCALER_LIST = {c.TYPE: c for c in (MultiplyRewardScaler, ClipRewardScaler, MinMaxRewardScaler,
                                          StandardRewardScaler, ReturnBasedRewardScaler)}


def _create(registry, what, name, **kwargs):
    assert name in registry, f"{name} seems not to be registered."
    return registry[name](**kwargs)


def create_scaler(name: str, **kwargs):
    """preprocessing/scalers.py:383-397."""
    return _create(SCALER_LIST, "scaler", name, **kwargs)


def create_action_scaler(name: str, **kwargs):
    """preprocessing/action_scalers.py:229-243."""
    return _create(ACTION_SCALER_LIST, "action scaler", name, **kwargs)


def create_reward_scaler(name: str, **kwargs):
    """preprocessing/reward_scalers.py:512-526."""
    return _create(REWARD_SCALER_LIST, "reward scaler", name, **kwargs)


def _check(value, registry, create):
    """argument_utility.check_scaler / check_action_scaler / check_reward_scaler: instance, registered name, or None."""
    if value is None or isinstance(value, _Scaler):
        return value
    if isinstance(value, str):
        return create(value)
    raise ValueError(f"unsupported scaler {value!r}")


def check_scaler(value):
    return _check(value, SCALER_LIST, create_scaler)


def check_action_scaler(value):
    return _check(value, ACTION_SCALER_LIST, create_action_scaler)


def check_reward_scaler(value):
    return _check(value, REWARD_SCALER_LIST, create_reward_scaler)
