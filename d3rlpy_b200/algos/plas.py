"""PLAS: same constructor/defaults as d3rlpy.algos.PLAS (d3rlpy/algos/plas.py:94-209)."""
from __future__ import annotations

from typing import Any, Dict

from .base import IMPL_NOT_INITIALIZED_ERROR, AlgoBase, _hidden_units
from .torch.plas_impl import PLASImpl


class PLAS(AlgoBase):
    IMPL = PLASImpl

    def __init__(self, *, actor_learning_rate: float = 1e-4, critic_learning_rate: float = 1e-3,
                 imitator_learning_rate: float = 1e-4, actor_optim_factory=None, critic_optim_factory=None,
                 imitator_optim_factory=None, actor_encoder_factory="default", critic_encoder_factory="default",
                 imitator_encoder_factory="default", q_func_factory="mean", batch_size: int = 100, n_frames: int = 1,
                 n_steps: int = 1, gamma: float = 0.99, tau: float = 0.005, n_critics: int = 2,
                 update_actor_interval: int = 1, lam: float = 0.75, warmup_steps: int = 500000, beta: float = 0.5,
                 use_gpu=0, scaler=None, action_scaler=None, reward_scaler=None, impl=None, seed: int = 0, **kwargs: Any):
        super().__init__(batch_size, n_frames, n_steps, gamma, scaler, action_scaler, reward_scaler, use_gpu, kwargs)
        if q_func_factory != "mean":
            raise ValueError("only the mean Q function is on the accelerated path")
        for f in (actor_optim_factory, critic_optim_factory, imitator_optim_factory):
            if f is not None:
                raise ValueError("only AdamFactory() defaults are on the accelerated path")
        self._actor_learning_rate, self._critic_learning_rate = actor_learning_rate, critic_learning_rate
        self._imitator_learning_rate = imitator_learning_rate
        self._actor_hidden = _hidden_units(actor_encoder_factory, [256, 256])
        self._critic_hidden = _hidden_units(critic_encoder_factory, [256, 256])
        self._imitator_hidden = _hidden_units(imitator_encoder_factory, [256, 256])
        self._tau, self._n_critics, self._update_actor_interval = tau, n_critics, update_actor_interval
        self._lam, self._warmup_steps, self._beta = lam, warmup_steps, beta
        self._impl, self._seed = impl, seed
        self._factories = {"actor_encoder_factory": actor_encoder_factory,
                           "critic_encoder_factory": critic_encoder_factory,
                           "imitator_encoder_factory": imitator_encoder_factory}

    def _create_impl(self, observation_shape, action_size) -> None:
        self._impl = self.IMPL(
            observation_shape=observation_shape, action_size=action_size,
            actor_learning_rate=self._actor_learning_rate, critic_learning_rate=self._critic_learning_rate,
            imitator_learning_rate=self._imitator_learning_rate, actor_hidden=self._actor_hidden,
            critic_hidden=self._critic_hidden, imitator_hidden=self._imitator_hidden, gamma=self._gamma, tau=self._tau,
            n_critics=self._n_critics, lam=self._lam, beta=self._beta, use_gpu=self._use_gpu, scaler=self._scaler,
            action_scaler=self._action_scaler, reward_scaler=self._reward_scaler, seed=self._seed, **self._kwargs)
        self._impl.build()

    def _flags(self):
        return self._grad_step < self._warmup_steps, self._grad_step % self._update_actor_interval == 0

    def _update(self, batch) -> Dict[str, float]:
        """plas.py:189-206: the VAE alone during warm-up, then critic every step and actor + both soft syncs every
        `update_actor_interval` steps (pre-increment grad_step)."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        return self._impl.update_fused(batch, *self._flags())

    def _update_async(self, batch):
        return self._impl.update_fused_async(batch, *self._flags())
