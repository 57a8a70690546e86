"""BEAR: same constructor/defaults as d3rlpy.algos.BEAR (d3rlpy/algos/bear.py:158-312)."""
from __future__ import annotations

from typing import Any, Dict

from .base import IMPL_NOT_INITIALIZED_ERROR, AlgoBase, _hidden_units
from .torch.bear_impl import BEARImpl


class BEAR(AlgoBase):
    IMPL = BEARImpl

    def __init__(self, *, actor_learning_rate: float = 1e-4, critic_learning_rate: float = 3e-4,
                 imitator_learning_rate: float = 3e-4, temp_learning_rate: float = 1e-4,
                 alpha_learning_rate: float = 1e-3, actor_optim_factory=None, critic_optim_factory=None,
                 imitator_optim_factory=None, temp_optim_factory=None, alpha_optim_factory=None,
                 actor_encoder_factory="default", critic_encoder_factory="default", imitator_encoder_factory="default",
                 q_func_factory="mean", batch_size: int = 256, n_frames: int = 1, n_steps: int = 1, gamma: float = 0.99,
                 tau: float = 0.005, n_critics: int = 2, initial_temperature: float = 1.0, initial_alpha: float = 1.0,
                 alpha_threshold: float = 0.05, lam: float = 0.75, n_action_samples: int = 100,
                 n_target_samples: int = 10, n_mmd_action_samples: int = 4, mmd_kernel: str = "laplacian",
                 mmd_sigma: float = 20.0, vae_kl_weight: float = 0.5, warmup_steps: int = 40000, use_gpu=0, scaler=None,
                 action_scaler=None, reward_scaler=None, impl=None, seed: int = 0, **kwargs: Any):
        super().__init__(batch_size, n_frames, n_steps, gamma, scaler, action_scaler, reward_scaler, use_gpu, kwargs)
        if q_func_factory != "mean":
            raise ValueError("only the mean Q function is on the accelerated path")
        for f in (actor_optim_factory, critic_optim_factory, imitator_optim_factory, temp_optim_factory,
                  alpha_optim_factory):
            if f is not None:
                raise ValueError("only AdamFactory() defaults are on the accelerated path")
        self._actor_learning_rate, self._critic_learning_rate = actor_learning_rate, critic_learning_rate
        self._imitator_learning_rate = imitator_learning_rate
        self._temp_learning_rate, self._alpha_learning_rate = temp_learning_rate, alpha_learning_rate
        self._actor_hidden = _hidden_units(actor_encoder_factory, [256, 256])
        self._critic_hidden = _hidden_units(critic_encoder_factory, [256, 256])
        self._imitator_hidden = _hidden_units(imitator_encoder_factory, [256, 256])
        self._tau, self._n_critics = tau, n_critics
        self._initial_temperature, self._initial_alpha, self._alpha_threshold = initial_temperature, initial_alpha, alpha_threshold
        self._lam, self._n_action_samples, self._n_target_samples = lam, n_action_samples, n_target_samples
        self._n_mmd_action_samples, self._mmd_kernel, self._mmd_sigma = n_mmd_action_samples, mmd_kernel, mmd_sigma
        self._vae_kl_weight, self._warmup_steps = vae_kl_weight, warmup_steps
        self._impl, self._seed = impl, seed
        self._factories = {"actor_encoder_factory": actor_encoder_factory,
                           "critic_encoder_factory": critic_encoder_factory,
                           "imitator_encoder_factory": imitator_encoder_factory}

    def _create_impl(self, observation_shape, action_size) -> None:
        self._impl = self.IMPL(
            observation_shape=observation_shape, action_size=action_size,
            actor_learning_rate=self._actor_learning_rate, critic_learning_rate=self._critic_learning_rate,
            imitator_learning_rate=self._imitator_learning_rate, temp_learning_rate=self._temp_learning_rate,
            alpha_learning_rate=self._alpha_learning_rate, actor_hidden=self._actor_hidden,
            critic_hidden=self._critic_hidden, imitator_hidden=self._imitator_hidden, gamma=self._gamma, tau=self._tau,
            n_critics=self._n_critics, initial_temperature=self._initial_temperature, initial_alpha=self._initial_alpha,
            alpha_threshold=self._alpha_threshold, lam=self._lam, n_action_samples=self._n_action_samples,
            n_target_samples=self._n_target_samples, n_mmd_action_samples=self._n_mmd_action_samples,
            mmd_kernel=self._mmd_kernel, mmd_sigma=self._mmd_sigma, vae_kl_weight=self._vae_kl_weight,
            use_gpu=self._use_gpu, scaler=self._scaler, action_scaler=self._action_scaler,
            reward_scaler=self._reward_scaler, seed=self._seed, **self._kwargs)
        self._impl.build()

    def _update(self, batch) -> Dict[str, float]:
        """bear.py:279-309; the warm-up test uses the pre-increment grad_step."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        return self._impl.update_fused(batch, self._grad_step < self._warmup_steps)

    def _update_async(self, batch):
        return self._impl.update_fused_async(batch, self._grad_step < self._warmup_steps)
