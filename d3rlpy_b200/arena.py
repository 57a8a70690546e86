"""Flat fp32 parameter arenas in HBM.

Every optimizer group (critic ensemble, actor, imitator, log_temp, log_alpha) lives in ONE
contiguous buffer with sibling buffers for grads / exp_avg / exp_avg_sq (and the Polyak target),
so Adam + soft_sync is a single HBM-bound pass (K10).  Named views reproduce the reference
modules' `state_dict()` keys (d3rlpy/torch_utility.py:97-110) so checkpoints and
`impl.q_function.q_funcs[i]`-style access keep working.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch


def _align4(n: int) -> int:
    return (n + 3) // 4 * 4


class ParamArena:
    def __init__(self, entries: Sequence[Tuple[str, Tuple[int, ...]]], members: int, device,
                 with_target: bool = False, member_key: str = "{name}", exports=None):
        """entries: per-member (name, shape) allocations.  member_key formats the state_dict key, e.g.
        "_q_funcs.{e}.{name}" for ensembles.  exports: optional [(key, parent, row0, nrows)] naming row
        slices of an allocation (concatenated heads) — they replace the parent in state_dict()."""
        self.members = members
        self.member_key = member_key
        self.offsets: Dict[str, int] = {}
        self.shapes: Dict[str, Tuple[int, ...]] = {}
        self.exports = list(exports) if exports else []
        off = 0
        for name, shape in entries:
            n = 1
            for s in shape:
                n *= s
            self.offsets[name] = off
            self.shapes[name] = tuple(shape)
            off = _align4(off + n)
        self.member_size = off
        self.size = off * members
        z = lambda: torch.zeros(self.size, dtype=torch.float32, device=device)
        self.params, self.grads, self.exp_avg, self.exp_avg_sq = z(), z(), z(), z()
        self.target = z() if with_target else None
        self.step = torch.zeros(1, dtype=torch.int32, device=device)  # Adam t (device-resident for graphs)

    # ---- raw addressing (bytes) used by the C-ABI calls
    def addr(self, which: str, name: str = None, member: int = 0) -> int:
        t = getattr(self, which)
        off = member * self.member_size + (self.offsets[name] if name else 0)
        return t.data_ptr() + 4 * off

    # ---- named tensor views
    def _exported(self):
        """[(key, alloc name, row0, nrows or None)] in registration order."""
        hidden = {p for _, p, _, _ in self.exports}
        out = [(n, n, 0, None) for n in self.offsets if n not in hidden]
        return out + [(k, p, r0, nr) for k, p, r0, nr in self.exports]

    def view(self, name: str, member: int = 0, which: str = "params") -> torch.Tensor:
        for k, parent, r0, nr in self.exports:
            if k == name:
                return self.view(parent, member, which)[r0:r0 + nr]
        t = getattr(self, which)
        off = member * self.member_size + self.offsets[name]
        shape = self.shapes[name]
        n = 1
        for s in shape:
            n *= s
        return t[off:off + n].view(shape)

    def state_dict(self, which: str = "params") -> "OrderedDict[str, torch.Tensor]":
        out = OrderedDict()
        for e in range(self.members):
            for key, _, _, _ in self._exported():
                out[self.member_key.format(e=e, name=key)] = self.view(key, e, which)
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor], which: str = "params") -> None:
        mine = self.state_dict(which)
        missing = [k for k in mine if k not in sd]
        if missing:
            raise KeyError(f"missing keys: {missing[:4]}...")
        with torch.no_grad():
            for k, v in mine.items():
                v.copy_(sd[k].to(v.device, torch.float32).reshape(v.shape))

    def sync_target_from_params(self) -> None:
        if self.target is not None:
            self.target.copy_(self.params)

    # Adam state in torch.optim.Adam.state_dict() layout (per named tensor)
    def optim_state(self) -> Dict[str, Dict[str, torch.Tensor]]:
        return {"step": self.step.clone(), "exp_avg": self.state_dict("exp_avg"),
                "exp_avg_sq": self.state_dict("exp_avg_sq")}
