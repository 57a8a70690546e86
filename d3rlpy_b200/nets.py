"""Dense networks over flat arenas: ReLU MLP trunk (+ narrow head) for all E ensemble members in
one launch per layer.  Mirrors VectorEncoder/VectorEncoderWithAction + `_fc`/`_mu`/`_logstd` heads
(d3rlpy/models/torch/encoders.py:236-339, q_functions/mean_q_function.py:61-72,
policies.py:47-59,82-97,127-181, imitators.py:13-72) without nn.Module or autograd: forward and the
hand-written backward are sequences of C-ABI kernel launches on the caller's stream.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

from ._lib import lib
from .arena import ParamArena


def _p(t) -> Optional[int]:
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


class DenseNet:
    """trunk: in_dim -> hidden[0] -> ... -> hidden[-1] (ReLU), head: hidden[-1] -> head_out."""

    def __init__(self, in_dim: int, hidden: Sequence[int], heads: Sequence[Tuple[str, int]], members: int,
                 device, trunk_prefix: str, member_key: str = "{name}", with_target: bool = False,
                 seed_gen: Optional[torch.Generator] = None):
        self.in_dim, self.hidden, self.members = in_dim, list(hidden), members
        self.heads = list(heads)
        self.head_out = sum(n for _, n in heads)
        self.trunk_prefix = trunk_prefix
        entries = []
        d = in_dim
        for i, h in enumerate(self.hidden):
            entries.append((f"{trunk_prefix}_fcs.{i}.weight", (h, d)))
            entries.append((f"{trunk_prefix}_fcs.{i}.bias", (h,)))
            d = h
        # all heads are ONE allocation (rows concatenated) so they run as one narrow GEMV; the
        # reference's per-head names are exported as row slices of it
        entries.append(("__head.weight", (self.head_out, d)))
        entries.append(("__head.bias", (self.head_out,)))
        exports, r0 = [], 0
        for name, n in self.heads:
            exports.append((f"{name}.weight", "__head.weight", r0, n))
            r0 += n
        r0 = 0
        for name, n in self.heads:
            exports.append((f"{name}.bias", "__head.bias", r0, n))
            r0 += n
        self.arena = ParamArena(entries, members, device, with_target=with_target, member_key=member_key,
                                exports=exports)
        self.feat = d
        self.device = device
        self._init_params(seed_gen)

    def _init_params(self, gen):
        """nn.Linear default init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias."""
        a = self.arena
        with torch.no_grad():
            for e in range(self.members):
                for name, shape in a.shapes.items():
                    fan_in = shape[1] if len(shape) == 2 else a.shapes[name[:-4] + "weight"][1]
                    bound = 1.0 / math.sqrt(fan_in)
                    v = (torch.rand(shape, generator=gen) * 2 - 1) * bound
                    a.view(name, e).copy_(v)
        a.sync_target_from_params()

    # ------------------------------------------------------------------ pointers
    def _w(self, which, i, member=0):
        return self.arena.addr(which, f"{self.trunk_prefix}_fcs.{i}.weight", member)

    def _b(self, which, i, member=0):
        return self.arena.addr(which, f"{self.trunk_prefix}_fcs.{i}.bias", member)

    def _hw(self, which, member=0):
        return self.arena.addr(which, "__head.weight", member)

    def _hb(self, which, member=0):
        return self.arena.addr(which, "__head.bias", member)

    def alloc_acts(self, rows: int, members: Optional[int] = None) -> List[torch.Tensor]:
        E = members or self.members
        return [torch.empty(E, rows, h, dtype=torch.float32, device=self.device) for h in self.hidden]

    # ------------------------------------------------------------------ forward
    def forward(self, which: str, x, ldx: int, stride_x: int, rows: int, acts: List[torch.Tensor], head_out,
                stream: int, head_tanh: bool = False, members: Optional[int] = None, member0: int = 0):
        """x: tensor or raw pointer, [rows, ldx] (stride_x=0: same input for every member).
        acts[l]: [E, rows, hidden[l]]; head_out: [E, rows, head_out]."""
        L = lib()
        E = members or self.members
        ms = self.arena.member_size
        cur, ld, sx = _p(x), ldx, stride_x
        d = self.in_dim
        for i, h in enumerate(self.hidden):
            y = acts[i]
            L.linear_forward(cur, ld, sx, self._w(which, i, member0), d, ms, self._b(which, i, member0), ms,
                             _p(y), h, rows * h, rows, h, d, E, 1, stream)
            cur, ld, sx, d = _p(y), h, rows * h, h
        if head_out is not None:
            n = self.head_out
            L.head_forward(cur, ld, sx, self._hw(which, member0), d, ms, self._hb(which, member0), ms,
                           _p(head_out), n, rows * n, rows, n, d, E, 1 if head_tanh else 0, stream)

    # ------------------------------------------------------------------ backward
    def backward(self, x, ldx: int, stride_x: int, rows: int, acts: List[torch.Tensor], d_head,
                 scratch: Tuple[torch.Tensor, torch.Tensor], stream: int, weight_grads: bool = True,
                 dx=None, lddx: int = 0, stride_dx: int = 0, dx_col0: int = 0, dx_cols: int = 0,
                 members: Optional[int] = None, member0: int = 0, d_head_ld: Optional[int] = None,
                 d_head_stride: Optional[int] = None):
        """d_head: [E, rows, head_out] gradient w.r.t. the (pre-activation) head output.
        Accumulates dW/db into arena.grads (RED) when weight_grads; optionally writes the gradient
        w.r.t. input columns [dx_col0, dx_col0+dx_cols) into dx."""
        L = lib()
        E = members or self.members
        ms = self.arena.member_size
        n = self.head_out
        feat = self.feat
        ldh = d_head_ld if d_head_ld is not None else n
        sdh = d_head_stride if d_head_stride is not None else rows * n
        last = acts[-1]
        if weight_grads:
            L.head_backward_weight(_p(d_head), ldh, sdh, _p(last), feat, rows * feat, self._hw("grads", member0), feat,
                                   ms, self._hb("grads", member0), ms, rows, n, feat, E, stream)
        dcur = scratch[0]
        L.head_backward_data(_p(d_head), ldh, sdh, self._hw("params", member0), feat, ms, _p(dcur), feat, rows * feat,
                             _p(last), feat, rows * feat, rows, n, feat, E, stream)
        which = 0
        for i in range(len(self.hidden) - 1, -1, -1):
            h = self.hidden[i]
            d_in = self.hidden[i - 1] if i > 0 else self.in_dim
            if i > 0:
                inp, ldi, si = _p(acts[i - 1]), d_in, rows * d_in
            else:
                inp, ldi, si = _p(x), ldx, stride_x
            if weight_grads:
                L.linear_backward_weight(_p(dcur), h, rows * h, inp, ldi, si, self._w("grads", i, member0), d_in, ms,
                                         self._b("grads", i, member0), ms, rows, h, d_in, E, stream)
            if i > 0:
                dnext = scratch[1 - which]
                L.linear_backward_data(_p(dcur), h, rows * h, self._w("params", i, member0), d_in, ms, _p(dnext), d_in,
                                       rows * d_in, inp, ldi, si, rows, h, d_in, E, stream)
                dcur, which = dnext, 1 - which
            elif dx is not None:
                L.linear_backward_data(_p(dcur), h, rows * h, self._w("params", 0, member0) + 4 * dx_col0, d_in, ms,
                                       _p(dx), lddx, stride_dx, None, 0, 0, rows, h, dx_cols, E, stream)

    def alloc_scratch(self, rows: int, members: Optional[int] = None):
        E = members or self.members
        hm = max(self.hidden)
        return (torch.empty(E, rows, hm, dtype=torch.float32, device=self.device),
                torch.empty(E, rows, hm, dtype=torch.float32, device=self.device))

    # ------------------------------------------------------------------ optimizer
    def adam(self, lr: float, stream: int, betas=(0.9, 0.999), eps=1e-8, tau: Optional[float] = None):
        a = self.arena
        lib().adam_step(_p(a.params), _p(a.grads), _p(a.exp_avg), _p(a.exp_avg_sq),
                        _p(a.target) if tau is not None else None, a.size, _p(a.step), lr, betas[0], betas[1], eps,
                        tau if tau is not None else 0.0, 1, stream)
