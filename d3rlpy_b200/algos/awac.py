"""AWAC: same constructor/defaults as d3rlpy.algos.AWAC (d3rlpy/algos/awac.py:98-191)."""
from __future__ import annotations

from typing import Any, Dict

from .base import IMPL_NOT_INITIALIZED_ERROR, AlgoBase, _hidden_units
from .torch.awr_impl import AWACImpl


def _adam_weight_decay(factory, default: float) -> float:
    """AdamFactory(weight_decay=...) is the only optimizer option on the accelerated path."""
    if factory is None:
        return default
    params = getattr(factory, "get_params", lambda: {})()
    wd = params.get("weight_decay", getattr(factory, "weight_decay", None))
    if wd is None:
        raise ValueError("only AdamFactory(weight_decay=...) is on the accelerated path")
    return float(wd)


class AWAC(AlgoBase):
    IMPL = AWACImpl
    WEIGHT_DECAY_OPTIMS = ("actor_optim_factory",)   # params.json carries the actor's weight decay inside its AdamFactory

    def __init__(self, *, actor_learning_rate: float = 3e-4, critic_learning_rate: float = 3e-4,
                 actor_optim_factory=None, critic_optim_factory=None, actor_encoder_factory="default",
                 critic_encoder_factory="default", q_func_factory="mean", batch_size: int = 1024, n_frames: int = 1,
                 n_steps: int = 1, gamma: float = 0.99, tau: float = 0.005, lam: float = 1.0, n_action_samples: int = 1,
                 n_critics: int = 2, update_actor_interval: int = 1, use_gpu=0, scaler=None, action_scaler=None,
                 reward_scaler=None, impl=None, seed: int = 0, **kwargs: Any):
        super().__init__(batch_size, n_frames, n_steps, gamma, scaler, action_scaler, reward_scaler, use_gpu, kwargs)
        if q_func_factory != "mean":
            raise ValueError("only the mean Q function is on the accelerated path")
        if critic_optim_factory is not None:
            raise ValueError("only AdamFactory() defaults are on the accelerated path")
        self._actor_weight_decay = _adam_weight_decay(actor_optim_factory, 1e-4)   # awac.py:105
        self._actor_learning_rate, self._critic_learning_rate = actor_learning_rate, critic_learning_rate
        self._actor_hidden = _hidden_units(actor_encoder_factory, [256, 256])
        self._critic_hidden = _hidden_units(critic_encoder_factory, [256, 256])
        self._tau, self._lam, self._n_action_samples = tau, lam, n_action_samples
        self._n_critics, self._update_actor_interval = n_critics, update_actor_interval
        self._impl, self._seed = impl, seed
        self._factories = {"actor_encoder_factory": actor_encoder_factory, "critic_encoder_factory": critic_encoder_factory}

    def _create_impl(self, observation_shape, action_size) -> None:
        self._impl = self.IMPL(
            observation_shape=observation_shape, action_size=action_size,
            actor_learning_rate=self._actor_learning_rate, critic_learning_rate=self._critic_learning_rate,
            actor_hidden=self._actor_hidden, critic_hidden=self._critic_hidden, gamma=self._gamma, tau=self._tau,
            lam=self._lam, n_action_samples=self._n_action_samples, n_critics=self._n_critics,
            actor_weight_decay=self._actor_weight_decay, use_gpu=self._use_gpu, scaler=self._scaler,
            action_scaler=self._action_scaler, reward_scaler=self._reward_scaler, seed=self._seed, **self._kwargs)
        self._impl.build()

    def _actor_step(self) -> bool:
        return self._grad_step % self._update_actor_interval == 0

    def _update(self, batch) -> Dict[str, float]:
        """awac.py:176-191 — critic every step; actor + both soft syncs every `update_actor_interval` steps, tested on
        the pre-increment grad_step."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        return self._impl.update_fused(batch, self._actor_step())

    def _update_async(self, batch):
        return self._impl.update_fused_async(batch, self._actor_step())
