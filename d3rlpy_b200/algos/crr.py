"""CRR: same constructor/defaults as d3rlpy.algos.CRR (d3rlpy/algos/crr.py:137-247)."""
from __future__ import annotations

from typing import Any, Dict

from .base import IMPL_NOT_INITIALIZED_ERROR, AlgoBase, _hidden_units
from .torch.awr_impl import CRRImpl


class CRR(AlgoBase):
    IMPL = CRRImpl

    def __init__(self, *, actor_learning_rate: float = 3e-4, critic_learning_rate: float = 3e-4,
                 actor_optim_factory=None, critic_optim_factory=None, actor_encoder_factory="default",
                 critic_encoder_factory="default", q_func_factory="mean", batch_size: int = 100, n_frames: int = 1,
                 n_steps: int = 1, gamma: float = 0.99, beta: float = 1.0, n_action_samples: int = 4,
                 advantage_type: str = "mean", weight_type: str = "exp", max_weight: float = 20.0, n_critics: int = 1,
                 target_update_type: str = "hard", tau: float = 5e-3, target_update_interval: int = 100,
                 update_actor_interval: int = 1, use_gpu=0, scaler=None, action_scaler=None, reward_scaler=None,
                 impl=None, seed: int = 0, **kwargs: Any):
        super().__init__(batch_size, n_frames, n_steps, gamma, scaler, action_scaler, reward_scaler, use_gpu, kwargs)
        if q_func_factory != "mean":
            raise ValueError("only the mean Q function is on the accelerated path")
        if actor_optim_factory is not None or critic_optim_factory is not None:
            raise ValueError("only AdamFactory() defaults are on the accelerated path")
        self._actor_learning_rate, self._critic_learning_rate = actor_learning_rate, critic_learning_rate
        self._actor_hidden = _hidden_units(actor_encoder_factory, [256, 256])
        self._critic_hidden = _hidden_units(critic_encoder_factory, [256, 256])
        self._beta, self._n_action_samples = beta, n_action_samples
        self._advantage_type, self._weight_type, self._max_weight = advantage_type, weight_type, max_weight
        self._n_critics, self._target_update_type, self._tau = n_critics, target_update_type, tau
        self._target_update_interval, self._update_actor_interval = target_update_interval, update_actor_interval
        self._impl, self._seed = impl, seed
        self._factories = {"actor_encoder_factory": actor_encoder_factory, "critic_encoder_factory": critic_encoder_factory}

    def _create_impl(self, observation_shape, action_size) -> None:
        self._impl = self.IMPL(
            observation_shape=observation_shape, action_size=action_size,
            actor_learning_rate=self._actor_learning_rate, critic_learning_rate=self._critic_learning_rate,
            actor_hidden=self._actor_hidden, critic_hidden=self._critic_hidden, gamma=self._gamma, tau=self._tau,
            beta=self._beta, n_action_samples=self._n_action_samples, advantage_type=self._advantage_type,
            weight_type=self._weight_type, max_weight=self._max_weight, n_critics=self._n_critics, use_gpu=self._use_gpu,
            scaler=self._scaler, action_scaler=self._action_scaler, reward_scaler=self._reward_scaler, seed=self._seed,
            **self._kwargs)
        self._impl.build()

    def _target_update(self) -> str:
        if self._target_update_type == "hard":
            return "hard" if self._grad_step % self._target_update_interval == 0 else "none"
        if self._target_update_type == "soft":
            return "soft"
        raise ValueError(f"invalid target_update_type: {self._target_update_type}")   # crr.py:240-243

    def _update(self, batch) -> Dict[str, float]:
        """crr.py:226-244: critic, actor, then hard targets every `target_update_interval` steps (pre-increment
        grad_step) or soft targets every step."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        return self._impl.update_fused(batch, self._target_update())

    def _update_async(self, batch):
        return self._impl.update_fused_async(batch, self._target_update())
