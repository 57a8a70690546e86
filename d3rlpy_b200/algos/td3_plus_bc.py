"""TD3PlusBC: same constructor/defaults as d3rlpy.algos.TD3PlusBC (d3rlpy/algos/td3_plus_bc.py:98-192)."""
from __future__ import annotations

from typing import Any, Dict

from .base import IMPL_NOT_INITIALIZED_ERROR, AlgoBase, _hidden_units
from .torch.td3_plus_bc_impl import TD3PlusBCImpl


class TD3PlusBC(AlgoBase):
    IMPL = TD3PlusBCImpl
    SUPPORTS_QR = True

    def __init__(self, *, actor_learning_rate: float = 3e-4, critic_learning_rate: float = 3e-4,
                 actor_optim_factory=None, critic_optim_factory=None, actor_encoder_factory="default",
                 critic_encoder_factory="default", q_func_factory="mean", batch_size: int = 256, n_frames: int = 1,
                 n_steps: int = 1, gamma: float = 0.99, tau: float = 0.005, n_critics: int = 2,
                 target_smoothing_sigma: float = 0.2, target_smoothing_clip: float = 0.5, alpha: float = 2.5,
                 update_actor_interval: int = 2, use_gpu=0, scaler="standard", action_scaler=None,
                 reward_scaler=None, impl=None, seed: int = 0, **kwargs: Any):
        if scaler == "standard":
            from ..preprocessing import StandardScaler

            scaler = StandardScaler()
        super().__init__(batch_size, n_frames, n_steps, gamma, scaler, action_scaler, reward_scaler, use_gpu, kwargs)
        from .dqn import _n_quantiles_of

        self._n_quantiles = _n_quantiles_of(q_func_factory)   # "mean", "qr" or QRQFunctionFactory(n_quantiles <= 32)
        if actor_optim_factory is not None or critic_optim_factory is not None:
            raise ValueError("only AdamFactory() defaults are on the accelerated path")
        self._actor_learning_rate, self._critic_learning_rate = actor_learning_rate, critic_learning_rate
        self._actor_hidden = _hidden_units(actor_encoder_factory, [256, 256])
        self._critic_hidden = _hidden_units(critic_encoder_factory, [256, 256])
        self._tau, self._n_critics = tau, n_critics
        self._target_smoothing_sigma, self._target_smoothing_clip = target_smoothing_sigma, target_smoothing_clip
        self._alpha, self._update_actor_interval = alpha, update_actor_interval
        self._impl, self._seed = impl, seed
        self._factories = {"actor_encoder_factory": actor_encoder_factory,
                           "critic_encoder_factory": critic_encoder_factory, "q_func_factory": q_func_factory}

    def _create_impl(self, observation_shape, action_size) -> None:
        self._impl = self.IMPL(
            observation_shape=observation_shape, action_size=action_size,
            actor_learning_rate=self._actor_learning_rate, critic_learning_rate=self._critic_learning_rate,
            actor_hidden=self._actor_hidden, critic_hidden=self._critic_hidden, gamma=self._gamma, tau=self._tau,
            n_critics=self._n_critics, target_smoothing_sigma=self._target_smoothing_sigma,
            target_smoothing_clip=self._target_smoothing_clip, alpha=self._alpha, use_gpu=self._use_gpu,
            scaler=self._scaler, action_scaler=self._action_scaler, reward_scaler=self._reward_scaler,
            seed=self._seed, n_quantiles=self._n_quantiles, **self._kwargs)
        self._impl.build()

    def _update(self, batch) -> Dict[str, float]:
        """td3_plus_bc.py:177-192 — actor/targets every `update_actor_interval` steps, tested on the
        pre-increment grad_step."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        actor_step = self._grad_step % self._update_actor_interval == 0
        return self._impl.update_fused(batch, actor_step)

    def _update_async(self, batch):
        return self._impl.update_fused_async(batch, self._grad_step % self._update_actor_interval == 0)
