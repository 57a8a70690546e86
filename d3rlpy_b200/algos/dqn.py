"""DQN / DoubleDQN / DiscreteCQL: same constructors/defaults as d3rlpy.algos.{DQN,DoubleDQN,DiscreteCQL}
(d3rlpy/algos/dqn.py:69-202, d3rlpy/algos/cql.py:317-372)."""
from __future__ import annotations

from typing import Any, Dict

from .base import IMPL_NOT_INITIALIZED_ERROR, AlgoBase, _hidden_units
from .torch.dqn_impl import DiscreteCQLImpl, DoubleDQNImpl, DQNImpl


class PixelEncoderFactory:
    """Same constructor as d3rlpy.models.encoders.PixelEncoderFactory (filters, feature_size)."""

    TYPE = "pixel"

    def __init__(self, filters=None, feature_size: int = 512, activation: str = "relu", use_batch_norm: bool = False,
                 dropout_rate=None):
        if activation != "relu" or use_batch_norm or dropout_rate is not None:
            raise ValueError("B200 path supports ReLU conv encoders without BN/dropout")
        self.filters = list(filters) if filters is not None else [(32, 8, 4), (64, 4, 2), (64, 3, 1)]
        self.feature_size = feature_size


class QRQFunctionFactory:
    """Same constructor as d3rlpy.models.q_functions.QRQFunctionFactory (models/q_functions.py:123-165)."""

    TYPE = "qr"

    def __init__(self, share_encoder: bool = False, n_quantiles: int = 32, **kwargs: Any):
        if share_encoder:
            raise ValueError("share_encoder is not on the accelerated path")
        self._share_encoder, self._n_quantiles = share_encoder, n_quantiles

    @property
    def n_quantiles(self) -> int:
        return self._n_quantiles

    def get_type(self) -> str:
        return self.TYPE

    def get_params(self, deep: bool = False) -> Dict[str, Any]:
        return {"share_encoder": self._share_encoder, "n_quantiles": self._n_quantiles}


def _n_quantiles_of(q_func_factory) -> int:
    """0 for the mean Q function, n_quantiles for "qr" / a QRQFunctionFactory (ours or the reference's)."""
    f = q_func_factory
    if f == "mean" or getattr(f, "TYPE", None) == "mean":
        return 0
    if f == "qr":
        return 32  # create_q_func_factory("qr"): QRQFunctionFactory() defaults
    if getattr(f, "TYPE", None) == "qr":
        if getattr(f, "share_encoder", False) or getattr(f, "_share_encoder", False):
            raise ValueError("share_encoder is not on the accelerated path")
        return int(f.n_quantiles)
    raise ValueError("the mean and quantile-regression (qr) Q functions are on the accelerated path; iqn/fqf are not")


class DQN(AlgoBase):
    IMPL = DQNImpl
    SUPPORTS_QR = True
    DISCRETE_ACTIONS = True    # get_action_type(): ActionSpace.DISCRETE (algos/dqn.py:134-135)

    def __init__(self, *, learning_rate: float = 6.25e-5, optim_factory=None, encoder_factory="default",
                 q_func_factory="mean", batch_size: int = 32, n_frames: int = 1, n_steps: int = 1, gamma: float = 0.99,
                 n_critics: int = 1, target_update_interval: int = 8000, use_gpu=0, scaler=None, reward_scaler=None,
                 impl=None, seed: int = 0, **kwargs: Any):
        if kwargs.pop("action_scaler", None) is not None:  # params.json of the reference carries "action_scaler": null
            raise ValueError("discrete algorithms take no action scaler")
        super().__init__(batch_size, n_frames, n_steps, gamma, scaler, None, reward_scaler, use_gpu, kwargs)
        self._n_quantiles = _n_quantiles_of(q_func_factory)
        if optim_factory is not None:
            raise ValueError("only AdamFactory() defaults are on the accelerated path")
        self._learning_rate, self._n_critics = learning_rate, n_critics
        self._target_update_interval = target_update_interval
        self._encoder_factory = encoder_factory
        self._impl, self._seed = impl, seed
        self._factories = {"encoder_factory": encoder_factory, "q_func_factory": q_func_factory}

    def _impl_kwargs(self) -> Dict[str, Any]:
        return {}

    def _create_impl(self, observation_shape, action_size) -> None:
        ef = self._encoder_factory
        kw = dict(self._kwargs)
        if len(observation_shape) == 3:
            if isinstance(ef, PixelEncoderFactory) or hasattr(ef, "feature_size") or hasattr(ef, "_feature_size"):
                kw["feature_size"] = getattr(ef, "feature_size", None) or getattr(ef, "_feature_size")
                kw["filters"] = getattr(ef, "filters", None) or getattr(ef, "_filters", None)
            hidden = []
        else:
            hidden = _hidden_units(ef, [256, 256])
        kw.update(self._impl_kwargs())
        kw["n_quantiles"] = self._n_quantiles
        self._impl = self.IMPL(observation_shape=observation_shape, action_size=action_size,
                               learning_rate=self._learning_rate, hidden=hidden, gamma=self._gamma,
                               n_critics=self._n_critics, use_gpu=self._use_gpu, scaler=self._scaler,
                               reward_scaler=self._reward_scaler, seed=self._seed, **kw)
        self._impl.build()

    def _update(self, batch) -> Dict[str, float]:
        """dqn.py:127-132: update, then hard target copy when grad_step % interval == 0 (pre-increment)."""
        assert self._impl is not None, IMPL_NOT_INITIALIZED_ERROR
        return self._impl.update_fused(batch, self._grad_step % self._target_update_interval == 0)

    def _update_async(self, batch):
        return self._impl.update_fused_async(batch, self._grad_step % self._target_update_interval == 0)


class DoubleDQN(DQN):
    IMPL = DoubleDQNImpl


class NFQ(DQN):
    """Neural Fitted Q Iteration (d3rlpy/algos/nfq.py:74-131): DQNImpl with the target network copied after EVERY
    update (`_update`: `impl.update(batch)`, `impl.update_target()`), i.e. DQN's schedule with an interval of one."""

    def __init__(self, **kw: Any):
        if "target_update_interval" in kw:
            raise TypeError("NFQ has no `target_update_interval`: the target is synchronised on every update")
        super().__init__(target_update_interval=1, **kw)

    def get_params(self, deep: bool = True):
        params = super().get_params(deep)
        params.pop("target_update_interval", None)
        return params


class DiscreteCQL(DoubleDQN):
    IMPL = DiscreteCQLImpl

    def __init__(self, *, alpha: float = 1.0, **kw: Any):
        super().__init__(**kw)
        self._alpha = alpha

    def _impl_kwargs(self):
        return {"alpha": self._alpha}
