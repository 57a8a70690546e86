"""DDPG: same constructor/defaults as d3rlpy.algos.DDPG (d3rlpy/algos/ddpg.py:96-180)."""
from __future__ import annotations

from typing import Any

from .td3 import TD3
from .torch.td3_plus_bc_impl import DDPGImpl

_TD3_ONLY = ("target_smoothing_sigma", "target_smoothing_clip", "update_actor_interval", "alpha")


class DDPG(TD3):
    """`_update` (ddpg.py:168-176): critic step, actor step, both soft syncs on EVERY update.  DDPGImpl.compute_target
    (algos/torch/ddpg_impl.py:275-284) is TD3's target without smoothing noise — `clamp(pi'(s'), -1, 1)` — so the
    update graph is TD3's with sigma = 0 and an actor interval of 1 (defaults: one critic, batch 100, no scaler)."""

    IMPL = DDPGImpl

    def __init__(self, *, batch_size: int = 100, n_critics: int = 1, scaler=None, **kw: Any):
        for k in _TD3_ONLY:
            if k in kw:
                raise TypeError(f"DDPG has no `{k}`")
        super().__init__(batch_size=batch_size, n_critics=n_critics, scaler=scaler, target_smoothing_sigma=0.0,
                         target_smoothing_clip=0.5, update_actor_interval=1, **kw)

    def get_params(self, deep: bool = True):
        params = super().get_params(deep)
        for k in _TD3_ONLY:
            params.pop(k, None)
        return params
