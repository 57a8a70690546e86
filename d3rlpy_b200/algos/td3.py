"""TD3: same constructor/defaults as d3rlpy.algos.TD3 (d3rlpy/algos/td3.py:90-176)."""
from __future__ import annotations

from typing import Any

from .td3_plus_bc import TD3PlusBC
from .torch.td3_plus_bc_impl import TD3Impl


class TD3(TD3PlusBC):
    """`_update` (td3.py:161-176) is TD3PlusBC's schedule: critic every step, actor + both soft syncs when the
    pre-increment grad_step is a multiple of `update_actor_interval`."""

    IMPL = TD3Impl

    def __init__(self, *, scaler=None, **kw: Any):
        if "alpha" in kw:
            raise TypeError("TD3 has no `alpha` (behaviour-cloning weight): use TD3PlusBC")
        super().__init__(scaler=scaler, **kw)

    def get_params(self, deep: bool = True):
        params = super().get_params(deep)
        params.pop("alpha", None)
        return params
