"""IQL: same constructor/defaults as d3rlpy.algos.IQL (d3rlpy/algos/iql.py:109-199)."""
from __future__ import annotations

from typing import Any, Dict

from .base import IMPL_NOT_INITIALIZED_ERROR, AlgoBase, _hidden_units
from .torch.iql_impl import IQLImpl


class IQL(AlgoBase):
    IMPL = IQLImpl
    HAS_Q_FUNC_FACTORY = False

    def __init__(self, *, actor_learning_rate: float = 3e-4, critic_learning_rate: float = 3e-4,
                 actor_optim_factory=None, critic_optim_factory=None, actor_encoder_factory="default",
                 critic_encoder_factory="default", value_encoder_factory="default", batch_size: int = 256,
                 n_frames: int = 1, n_steps: int = 1, gamma: float = 0.99, tau: float = 0.005, n_critics: int = 2,
                 expectile: float = 0.7, weight_temp: float = 3.0, max_weight: float = 100.0, use_gpu=0, scaler=None,
                 action_scaler=None, reward_scaler=None, impl=None, seed: int = 0, **kwargs: Any):
        super().__init__(batch_size, n_frames, n_steps, gamma, scaler, action_scaler, reward_scaler, use_gpu, kwargs)
        if actor_optim_factory is not None or critic_optim_factory is not None:
            raise ValueError("only AdamFactory() defaults are on the accelerated path")
        self._actor_learning_rate, self._critic_learning_rate = actor_learning_rate, critic_learning_rate
        self._actor_hidden = _hidden_units(actor_encoder_factory, [256, 256])
        self._critic_hidden = _hidden_units(critic_encoder_factory, [256, 256])
        self._value_hidden = _hidden_units(value_encoder_factory, [256, 256])
        self._tau, self._n_critics = tau, n_critics
        self._expectile, self._weight_temp, self._max_weight = expectile, weight_temp, max_weight
        self._impl, self._seed = impl, seed
        self._factories = {"actor_encoder_factory": actor_encoder_factory,
                           "critic_encoder_factory": critic_encoder_factory,
                           "value_encoder_factory": value_encoder_factory}

    def _create_impl(self, observation_shape, action_size) -> None:
        self._impl = self.IMPL(
            observation_shape=observation_shape, action_size=action_size,
            actor_learning_rate=self._actor_learning_rate, critic_learning_rate=self._critic_learning_rate,
            actor_hidden=self._actor_hidden, critic_hidden=self._critic_hidden, value_hidden=self._value_hidden,
            gamma=self._gamma, tau=self._tau, n_critics=self._n_critics, expectile=self._expectile,
            weight_temp=self._weight_temp, max_weight=self._max_weight, use_gpu=self._use_gpu, scaler=self._scaler,
            action_scaler=self._action_scaler, reward_scaler=self._reward_scaler, seed=self._seed, **self._kwargs)
        self._impl.build()

    def _update(self, batch) -> Dict[str, float]:
        """iql.py:186-199: critic (+ value) step, actor step, critic target sync — every update."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        return self._impl.update_fused(batch)

    def _update_async(self, batch):
        return self._impl.update_fused_async(batch)
