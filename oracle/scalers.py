"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's observation / action / reward scalers, the
transforms `TorchMiniBatch.__init__` applies to a minibatch (d3rlpy/torch_utility.py:179-185) and `predict*` apply
around the policy (d3rlpy/algos/torch/base.py:52-80).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module.

Each class restates `fit(transitions)` over flat numpy arrays of the transitions (observation / action / reward of
every transition, in dataset order) and `transform` / `reverse_transform` in float32 torch arithmetic with the
reference's operator order.  Pinned by tests/golden/scalers.npz (parameters fitted by, and tensors transformed by, the
unmodified reference: tests/golden/make_golden_scalers.py).
"""
from __future__ import annotations

import numpy as np
import torch


def _t(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32)


class MinMaxScaler:
    """preprocessing/scalers.py:119-254: running np.minimum / np.maximum over transition observations (:179-194),
    `(x - min) / (max - min)` on float32 tensors (:209-218)."""

    def __init__(self, minimum=None, maximum=None):
        self.minimum, self.maximum = minimum, maximum

    def fit(self, observations: np.ndarray):
        if self.minimum is None or self.maximum is None:
            obs = np.asarray(observations)
            self.minimum = obs.min(axis=0).reshape((1,) + obs.shape[1:])
            self.maximum = obs.max(axis=0).reshape((1,) + obs.shape[1:])
        return self

    def __call__(self, x):
        mn, mx = _t(self.minimum), _t(self.maximum)
        return (x - mn) / (mx - mn)

    def reverse(self, x):
        mn, mx = _t(self.minimum), _t(self.maximum)
        return ((mx - mn) * x) + mn


class StandardScaler:
    """preprocessing/scalers.py:256-366: mean = sum / count, std = sqrt(sum((x - mean)^2) / count) accumulated over
    transitions in the observation dtype (:318-343); `(x - mean) / (std + eps)` (:345-354)."""

    def __init__(self, mean=None, std=None, eps=1e-3):
        self.mean, self.std, self.eps = mean, std, eps

    def fit(self, observations: np.ndarray):
        if self.mean is None or self.std is None:
            obs = np.asarray(observations)
            total_sum = np.zeros(obs.shape[1:])
            for o in obs:
                total_sum += o
            mean = total_sum / obs.shape[0]
            total_sqsum = np.zeros(obs.shape[1:])
            for o in obs:
                total_sqsum += (o - mean.reshape(o.shape)) ** 2
            self.mean = mean.reshape((1,) + obs.shape[1:])
            self.std = np.sqrt(total_sqsum / obs.shape[0]).reshape((1,) + obs.shape[1:])
        return self

    def __call__(self, x):
        return (x - _t(self.mean)) / (_t(self.std) + self.eps)


class MinMaxActionScaler:
    """preprocessing/action_scalers.py:139-212: `((a - min) / (max - min)) * 2 - 1`, reverse
    `((max - min) * ((a + 1) / 2)) + min`."""

    def __init__(self, minimum=None, maximum=None):
        self.minimum, self.maximum = minimum, maximum

    def fit(self, actions: np.ndarray):
        if self.minimum is None or self.maximum is None:
            act = np.asarray(actions)
            self.minimum = act.min(axis=0).reshape((1,) + act.shape[1:])
            self.maximum = act.max(axis=0).reshape((1,) + act.shape[1:])
        return self

    def __call__(self, a):
        mn, mx = _t(self.minimum), _t(self.maximum)
        return ((a - mn) / (mx - mn)) * 2.0 - 1.0

    def reverse(self, a):
        mn, mx = _t(self.minimum), _t(self.maximum)
        return ((mx - mn) * ((a + 1.0) / 2.0)) + mn


class MultiplyRewardScaler:
    """preprocessing/reward_scalers.py:96-135."""

    def __init__(self, multiplier=1.0):
        self.multiplier = multiplier

    def fit(self, rewards, episode_returns=None):
        return self

    def __call__(self, r):
        return self.multiplier * r


class ClipRewardScaler:
    """preprocessing/reward_scalers.py:138-190: `multiplier * clamp(r, low, high)`."""

    def __init__(self, low=None, high=None, multiplier=1.0):
        self.low, self.high, self.multiplier = low, high, multiplier

    def fit(self, rewards, episode_returns=None):
        return self

    def __call__(self, r):
        return self.multiplier * r.clamp(self.low, self.high)


class MinMaxRewardScaler:
    """preprocessing/reward_scalers.py:193-281: python-float min / max of the transition rewards,
    `multiplier * (r - min) / (max - min)`."""

    def __init__(self, minimum=None, maximum=None, multiplier=1.0):
        self.minimum, self.maximum, self.multiplier = minimum, maximum, multiplier

    def fit(self, rewards, episode_returns=None):
        if self.minimum is None or self.maximum is None:
            rewards = [float(r) for r in np.asarray(rewards).reshape(-1)]
            self.minimum, self.maximum = float(np.min(rewards)), float(np.max(rewards))
        return self

    def __call__(self, r):
        base = self.maximum - self.minimum
        return self.multiplier * (r - self.minimum) / base


class StandardRewardScaler:
    """preprocessing/reward_scalers.py:284-376: float64 mean / population std of the transition rewards,
    `multiplier * (r - mean) / (std + eps)`."""

    def __init__(self, mean=None, std=None, eps=1e-3, multiplier=1.0):
        self.mean, self.std, self.eps, self.multiplier = mean, std, eps, multiplier

    def fit(self, rewards, episode_returns=None):
        if self.mean is None or self.std is None:
            rewards = [float(r) for r in np.asarray(rewards).reshape(-1)]
            self.mean, self.std = float(np.mean(rewards)), float(np.std(rewards))
        return self

    def __call__(self, r):
        nonzero_std = self.std + self.eps
        return self.multiplier * (r - self.mean) / nonzero_std


class ReturnBasedRewardScaler:
    """preprocessing/reward_scalers.py:379-488: max / min over the per-episode sums of transition rewards (python
    float accumulation in transition order), `multiplier * r / (return_max - return_min)`."""

    def __init__(self, return_max=None, return_min=None, multiplier=1.0):
        self.return_max, self.return_min, self.multiplier = return_max, return_min, multiplier

    def fit(self, rewards, episode_returns=None):
        if self.return_max is None or self.return_min is None:
            self.return_max, self.return_min = float(np.max(episode_returns)), float(np.min(episode_returns))
        return self

    def __call__(self, r):
        return self.multiplier * r / (self.return_max - self.return_min)


def episode_returns(rewards: np.ndarray, episode_of_transition: np.ndarray):
    """Per-episode sums as ReturnBasedRewardScaler.fit accumulates them (:452-463): `ret = 0.0; ret += reward` walking
    each episode's transitions first to last -- float64 accumulation of float32 rewards."""
    out = {}
    for r, e in zip(np.asarray(rewards).reshape(-1), np.asarray(episode_of_transition).reshape(-1)):
        out[int(e)] = out.get(int(e), 0.0) + float(r)
    return [out[k] for k in sorted(out)]
