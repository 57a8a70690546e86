"""BCQ: same constructor/defaults as d3rlpy.algos.BCQ (d3rlpy/algos/bcq.py:168-279)."""
from __future__ import annotations

from typing import Any, Dict

from .base import IMPL_NOT_INITIALIZED_ERROR, AlgoBase, _hidden_units
from .torch.bcq_impl import BCQImpl


class BCQ(AlgoBase):
    def __init__(self, *, actor_learning_rate: float = 1e-3, critic_learning_rate: float = 1e-3,
                 imitator_learning_rate: float = 1e-3, actor_optim_factory=None, critic_optim_factory=None,
                 imitator_optim_factory=None, actor_encoder_factory="default", critic_encoder_factory="default",
                 imitator_encoder_factory="default", q_func_factory="mean", batch_size: int = 100, n_frames: int = 1,
                 n_steps: int = 1, gamma: float = 0.99, tau: float = 0.005, n_critics: int = 2,
                 update_actor_interval: int = 1, lam: float = 0.75, n_action_samples: int = 100,
                 action_flexibility: float = 0.05, rl_start_step: int = 0, beta: float = 0.5, use_gpu=0, scaler=None,
                 action_scaler=None, reward_scaler=None, impl=None, seed: int = 0, **kwargs: Any):
        super().__init__(batch_size, n_frames, n_steps, gamma, scaler, action_scaler, reward_scaler, use_gpu, kwargs)
        if q_func_factory != "mean":
            raise ValueError("only the mean Q function is on the accelerated path")
        for f in (actor_optim_factory, critic_optim_factory, imitator_optim_factory):
            if f is not None:
                raise ValueError("only AdamFactory() defaults are on the accelerated path")
        self._actor_learning_rate, self._critic_learning_rate = actor_learning_rate, critic_learning_rate
        self._imitator_learning_rate = imitator_learning_rate
        # "default" vector encoders are 256x256 (models/encoders.py:DefaultEncoderFactory); the BCQ
        # reproduction passes [400,300] / [750,750] explicitly (reproductions/offline/bcq.py:21-34)
        self._actor_hidden = _hidden_units(actor_encoder_factory, [256, 256])
        self._critic_hidden = _hidden_units(critic_encoder_factory, [256, 256])
        self._imitator_hidden = _hidden_units(imitator_encoder_factory, [256, 256])
        self._tau, self._n_critics, self._update_actor_interval = tau, n_critics, update_actor_interval
        self._lam, self._n_action_samples, self._action_flexibility = lam, n_action_samples, action_flexibility
        self._rl_start_step, self._beta = rl_start_step, beta
        self._impl, self._seed = impl, seed
        self._factories = {"actor_encoder_factory": actor_encoder_factory, "critic_encoder_factory": critic_encoder_factory,
                           "imitator_encoder_factory": imitator_encoder_factory}

    def _create_impl(self, observation_shape, action_size) -> None:
        self._impl = BCQImpl(
            observation_shape=observation_shape, action_size=action_size,
            actor_learning_rate=self._actor_learning_rate, critic_learning_rate=self._critic_learning_rate,
            imitator_learning_rate=self._imitator_learning_rate, actor_hidden=self._actor_hidden,
            critic_hidden=self._critic_hidden, imitator_hidden=self._imitator_hidden, gamma=self._gamma,
            tau=self._tau, n_critics=self._n_critics, lam=self._lam, n_action_samples=self._n_action_samples,
            action_flexibility=self._action_flexibility, beta=self._beta, use_gpu=self._use_gpu, scaler=self._scaler,
            action_scaler=self._action_scaler, reward_scaler=self._reward_scaler, seed=self._seed, **self._kwargs)
        self._impl.build()

    def _update(self, batch) -> Dict[str, float]:
        """bcq.py:261-279: imitator always; critic once grad_step >= rl_start_step; actor + both soft
        syncs every update_actor_interval steps (pre-increment grad_step)."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        rl_step = self._grad_step >= self._rl_start_step
        actor_step = self._grad_step % self._update_actor_interval == 0
        return self._impl.update_fused(batch, rl_step, actor_step)

    def _update_async(self, batch):
        return self._impl.update_fused_async(batch, self._grad_step >= self._rl_start_step,
                                             self._grad_step % self._update_actor_interval == 0)
