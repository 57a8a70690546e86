"""`train_single_env` — the online training loop of d3rlpy/online/iterators.py:99-287 around the accelerated pieces:
actions from the CUDA evaluation path, transitions into the HBM `ReplayBuffer`, `algo.update` on the minibatch the
gather kernels assemble on the device.  Loop order, update condition (`total_step > update_start_step and
len(buffer) > batch_size`, every `update_interval` steps), TimeLimit handling and the final `clip_episode` are the
reference's; logging is reduced to the per-epoch metric means it would have committed (returned as a list)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import numpy as np


class StackedObservation:
    """preprocessing/stack.py:6-63: channel-wise ring of the last n_frames images (zeros after `clear`)."""

    def __init__(self, observation_shape, n_frames: int, dtype=np.uint8):
        self._c, self._n = observation_shape[0], n_frames
        self._stack = np.zeros((self._c * n_frames,) + tuple(observation_shape[1:]), dtype)

    def append(self, image: np.ndarray) -> None:
        assert image.dtype == self._stack.dtype
        self._stack = np.roll(self._stack, -self._c, axis=0)
        self._stack[self._c * (self._n - 1):] = image.copy()

    def eval(self) -> np.ndarray:
        return self._stack

    def clear(self) -> None:
        self._stack.fill(0)


def _setup_algo(algo, env) -> None:
    """iterators.py:76-96: observation / action scalers take their bounds from the environment's spaces, then the
    impl is built."""
    if algo.scaler:
        algo.scaler.fit_with_env(env)
    if algo.action_scaler:
        algo.action_scaler.fit_with_env(env)
    if algo.impl is None:
        algo.build_with_env(env)


def train_single_env(algo, env, buffer, explorer=None, n_steps: int = 1000000, n_steps_per_epoch: int = 10000,
                     update_interval: int = 1, update_start_step: int = 0, random_steps: int = 0,
                     timelimit_aware: bool = True,
                     callback: Optional[Callable[[Any, int, int], None]] = None) -> List[Dict[str, float]]:
    _setup_algo(algo, env)
    observation_shape = env.observation_space.shape
    is_image = len(observation_shape) == 3
    stacked_frame = StackedObservation(observation_shape, algo.n_frames) if is_image else None
    history: List[Dict[str, float]] = []
    sums: Dict[str, List[float]] = {}

    def add_metric(name, value):
        sums.setdefault(name, []).append(float(value))

    observation = env.reset()
    rollout_return = 0.0
    for total_step in range(1, n_steps + 1):
        if is_image:
            stacked_frame.append(observation)
            fed_observation = stacked_frame.eval()
        else:
            observation = observation.astype("f4")
            fed_observation = observation
        if total_step < random_steps:
            action = env.action_space.sample()
        elif explorer:
            x = fed_observation.reshape((1,) + fed_observation.shape)
            action = explorer.sample(algo, x, total_step)[0]
        else:
            action = algo.sample_action(fed_observation[None])[0]
        next_observation, reward, terminal, info = env.step(action)
        rollout_return += reward
        if timelimit_aware and "TimeLimit.truncated" in info:
            clip_episode, terminal = True, False
        else:
            clip_episode = terminal
        buffer.append(observation=observation, action=action, reward=reward, terminal=terminal,
                      clip_episode=clip_episode)
        if clip_episode:
            observation = env.reset()
            add_metric("rollout_return", rollout_return)
            rollout_return = 0.0
            if is_image:
                stacked_frame.clear()
        else:
            observation = next_observation
        epoch = total_step // n_steps_per_epoch
        if total_step > update_start_step and len(buffer) > algo.batch_size:
            if total_step % update_interval == 0:
                batch = buffer.sample(batch_size=algo.batch_size, n_frames=algo.n_frames, n_steps=algo.n_steps,
                                      gamma=algo.gamma)
                for name, val in algo.update(batch).items():
                    add_metric(name, val)
        if callback:
            callback(algo, epoch, total_step)
        if epoch > 0 and total_step % n_steps_per_epoch == 0:
            history.append({k: float(np.mean(v)) for k, v in sums.items()})
            sums = {}
    buffer.clip_episode()
    return history
