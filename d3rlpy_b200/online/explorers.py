"""Exploration wrappers with the constructors and `sample(algo, x, step)` of d3rlpy.online.explorers
(d3rlpy/online/explorers.py:28-171).  Host-side numpy on top of `algo.predict` (the greedy action comes from the CUDA
evaluation path); they consume the global numpy stream in the reference's order."""
from __future__ import annotations

import numpy as np


class Explorer:
    def sample(self, algo, x: np.ndarray, step: int) -> np.ndarray:
        raise NotImplementedError


class ConstantEpsilonGreedy(Explorer):
    """explorers.py:36-62."""

    def __init__(self, epsilon: float):
        self._epsilon = epsilon

    def sample(self, algo, x: np.ndarray, step: int) -> np.ndarray:
        greedy_actions = algo.predict(x)
        random_actions = np.random.randint(algo.action_size, size=x.shape[0])
        is_random = np.random.random(x.shape[0]) < self._epsilon
        return np.where(is_random, random_actions, greedy_actions)


class LinearDecayEpsilonGreedy(Explorer):
    """explorers.py:65-125."""

    def __init__(self, start_epsilon: float = 1.0, end_epsilon: float = 0.1, duration: int = 1000000):
        self._start_epsilon, self._end_epsilon, self._duration = start_epsilon, end_epsilon, duration

    def sample(self, algo, x: np.ndarray, step: int) -> np.ndarray:
        greedy_actions = algo.predict(x)
        random_actions = np.random.randint(algo.action_size, size=x.shape[0])
        is_random = np.random.random(x.shape[0]) < self.compute_epsilon(step)
        return np.where(is_random, random_actions, greedy_actions)

    def compute_epsilon(self, step: int) -> float:
        if step >= self._duration:
            return self._end_epsilon
        base = self._start_epsilon - self._end_epsilon
        return base * (1.0 - step / self._duration) + self._end_epsilon


class NormalNoise(Explorer):
    """explorers.py:128-171: ONE scalar draw added to every action component (the reference's documented quirk);
    the result is clipped to the MinMaxActionScaler's [minimum, maximum] when the algorithm carries one (`predict`
    returns actions already mapped back to that range), else to [-1, 1]."""

    def __init__(self, mean: float = 0.0, std: float = 0.1):
        self._mean, self._std = mean, std

    def sample(self, algo, x: np.ndarray, step: int) -> np.ndarray:
        action = algo.predict(x)
        noise = np.random.normal(self._mean, self._std)
        from ..preprocessing import MinMaxActionScaler

        scaler = getattr(algo, "action_scaler", None)
        if isinstance(scaler, MinMaxActionScaler):
            params = scaler.get_params()
            minimum, maximum = params["minimum"], params["maximum"]
        else:
            minimum, maximum = -1.0, 1.0
        return np.clip(action + noise, minimum, maximum)
