"""SAC: same constructor/defaults as d3rlpy.algos.SAC (d3rlpy/algos/sac.py:95-198)."""
from __future__ import annotations

from typing import Any, Dict

from .base import IMPL_NOT_INITIALIZED_ERROR, AlgoBase, _hidden_units
from .torch.sac_impl import SACImpl


class SAC(AlgoBase):
    def __init__(self, *, actor_learning_rate: float = 3e-4, critic_learning_rate: float = 3e-4,
                 temp_learning_rate: float = 3e-4, actor_optim_factory=None, critic_optim_factory=None,
                 temp_optim_factory=None, actor_encoder_factory="default", critic_encoder_factory="default",
                 q_func_factory="mean", batch_size: int = 256, n_frames: int = 1, n_steps: int = 1,
                 gamma: float = 0.99, tau: float = 0.005, n_critics: int = 2, initial_temperature: float = 1.0,
                 use_gpu=0, scaler=None, action_scaler=None, reward_scaler=None, impl=None, seed: int = 0,
                 **kwargs: Any):
        super().__init__(batch_size, n_frames, n_steps, gamma, scaler, action_scaler, reward_scaler, use_gpu, kwargs)
        if q_func_factory != "mean":
            raise ValueError("only the mean Q function is on the accelerated path")
        for f in (actor_optim_factory, critic_optim_factory, temp_optim_factory):
            if f is not None:
                raise ValueError("only AdamFactory() defaults are on the accelerated path")
        self._actor_learning_rate, self._critic_learning_rate = actor_learning_rate, critic_learning_rate
        self._temp_learning_rate = temp_learning_rate
        self._actor_hidden = _hidden_units(actor_encoder_factory, [256, 256])
        self._critic_hidden = _hidden_units(critic_encoder_factory, [256, 256])
        self._tau, self._n_critics, self._initial_temperature = tau, n_critics, initial_temperature
        self._impl, self._seed = impl, seed
        self._factories = {"actor_encoder_factory": actor_encoder_factory, "critic_encoder_factory": critic_encoder_factory}

    def _create_impl(self, observation_shape, action_size) -> None:
        self._impl = SACImpl(
            observation_shape=observation_shape, action_size=action_size,
            actor_learning_rate=self._actor_learning_rate, critic_learning_rate=self._critic_learning_rate,
            temp_learning_rate=self._temp_learning_rate, actor_hidden=self._actor_hidden,
            critic_hidden=self._critic_hidden, gamma=self._gamma, tau=self._tau, n_critics=self._n_critics,
            initial_temperature=self._initial_temperature, use_gpu=self._use_gpu, scaler=self._scaler,
            action_scaler=self._action_scaler, reward_scaler=self._reward_scaler, seed=self._seed, **self._kwargs)
        self._impl.build()

    def _update(self, batch) -> Dict[str, float]:
        """sac.py:177-198: temp -> critic -> actor -> both soft syncs, as one captured graph."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        return self._impl.update_fused(batch)

    def _update_async(self, batch):
        return self._impl.update_fused_async(batch)
