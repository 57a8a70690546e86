"""`soft_sync` / `hard_sync` with the reference's signatures (d3rlpy/torch_utility.py:27-41), running
as one fused pass over the flat arenas."""
from __future__ import annotations

import torch

from ._lib import lib


def _arena_of(view):
    net = getattr(view, "_net", None)
    if net is None:
        raise TypeError("expected impl.q_function / impl.policy style views of a d3rlpy_b200 network")
    return net.arena


def soft_sync(targ_model, model, tau: float) -> None:
    a = _arena_of(model)
    assert _arena_of(targ_model) is a and a.target is not None
    lib().soft_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, float(tau),
                    torch.cuda.current_stream(a.params.device).cuda_stream)


def hard_sync(targ_model, model) -> None:
    a = _arena_of(model)
    assert _arena_of(targ_model) is a and a.target is not None
    lib().hard_sync(a.target.data_ptr(), a.params.data_ptr(), a.size,
                    torch.cuda.current_stream(a.params.device).cuda_stream)


def soft_sync_tensors(target: torch.Tensor, params: torch.Tensor, tau: float) -> None:
    """Flat-tensor form used by tests: target = target*(1-tau) + tau*params (two roundings)."""
    assert target.is_cuda and target.dtype == torch.float32 and target.numel() == params.numel()
    lib().soft_sync(target.data_ptr(), params.data_ptr(), target.numel(), float(tau),
                    torch.cuda.current_stream(target.device).cuda_stream)
