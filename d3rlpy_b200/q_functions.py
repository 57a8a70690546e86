"""Callable ensemble Q functions over the CUDA forward: the module-level API of
d3rlpy/models/torch/q_functions/ensemble_q_function.py:69-184 that `impl.q_function` / `impl.targ_q_function` expose —
`__call__(x[, action], reduction)`, `compute_target(x, action, reduction, lam)`, `compute_error(...)`, `q_funcs[i]`,
plus `state_dict / load_state_dict / parameters` over the flat arenas under the reference's key names.

Inputs are the tensors the reference modules see (observations / actions already scaled); any device, copied to the
impl's device.  The member loop + `torch.cat` of the reference is ONE batched launch per layer (or one fused launch in
bf16 mode) for all members; `_reduce_ensemble` (min / max / mean / none / mix) is `d3b_ensemble_reduce`, the TD error
`d3b_td_error`.  Results are fresh device tensors.  Quantile-regression critics are not exposed through this API
(`predict_value` and the update path handle them)."""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Optional

import numpy as np
import torch

_MODES = {"min": 0, "max": 1, "mean": 2, "mix": 3}


class _ModuleView:
    """Stand-in for the nn.Module the reference exposes through `impl.q_function` / `impl.policy`:
    `state_dict()`, `load_state_dict()`, `parameters()` over arena views."""

    def __init__(self, net, which: str = "params"):
        self._net, self._which = net, which

    def state_dict(self):
        return self._net.arena.state_dict(self._which)

    def load_state_dict(self, sd):
        self._net.arena.load_state_dict(sd, self._which)
        st = torch.cuda.current_stream(self._net.device)
        self._net.refresh_shadow(self._which, st.cuda_stream)
        st.synchronize()

    def parameters(self):
        return list(self.state_dict().values())


def _dev_f32(impl, t) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t))
    with torch.cuda.stream(impl._stream_obj):
        return t.detach().to(impl._device, torch.float32).contiguous()


class _EnsembleBase(_ModuleView):
    def __init__(self, impl, which: str = "params", member0: int = 0, members: Optional[int] = None):
        super().__init__(impl._q_func, which)
        self._impl, self._member0 = impl, member0
        self._members = members if members is not None else impl._n_critics
        if getattr(impl, "_n_quantiles", 0):
            self._check = self._no_qr

    def _no_qr(self):
        raise NotImplementedError("the callable Q-function API covers mean Q functions; quantile-regression critics go "
                                  "through predict_value / the update path")

    def _check(self):
        return None

    @property
    def q_funcs(self) -> List["_EnsembleBase"]:
        """One single-member view per critic (ensemble_q_function.py:131-133)."""
        return [type(self)(self._impl, self._which, self._member0 + e, 1) for e in range(self._members)]

    def state_dict(self):
        sd = super().state_dict()
        if self._members == self._impl._n_critics and self._member0 == 0:
            return sd
        pre = f"_q_funcs.{self._member0}."   # a member view carries the member module's own keys
        return OrderedDict((k[len(pre):], v) for k, v in sd.items() if k.startswith(pre))

    def load_state_dict(self, sd):
        if self._members == self._impl._n_critics and self._member0 == 0:
            return super().load_state_dict(sd)
        full = super().state_dict()
        pre = f"_q_funcs.{self._member0}."
        with torch.no_grad():
            for k, v in sd.items():
                full[pre + k].copy_(torch.as_tensor(v).to(full[pre + k].device, torch.float32).reshape(full[pre + k].shape))
        st = torch.cuda.current_stream(self._net.device)
        self._net.refresh_shadow(self._which, st.cuda_stream)
        st.synchronize()

    # ---- shared pieces
    def _reduce(self, q: torch.Tensor, cols: int, reduction: str, lam: float) -> torch.Tensor:
        """q: [E, n * cols] contiguous device values -> _reduce_ensemble over the member axis."""
        impl, E = self._impl, self._members
        n = q.shape[1] // cols
        if reduction == "none":
            out = q.clone().view(E, n, cols)
        elif reduction in _MODES:
            out = torch.empty(n, cols, dtype=torch.float32, device=impl._device)
            impl._lib.ensemble_reduce(q.data_ptr(), q.shape[1], n * cols, E, _MODES[reduction], float(lam), out.data_ptr(),
                                      impl._stream)
        else:
            raise ValueError(f"invalid reduction: {reduction}")   # ensemble_q_function.py:24
        impl.sync()
        return out

    def _td(self, q: torch.Tensor, rewards, target, terminals, gamma, huber: bool) -> torch.Tensor:
        impl, E, n = self._impl, self._members, q.shape[1]
        tgt = _dev_f32(impl, target)
        assert tgt.ndim == 2, "target must be (batch, 1)"   # ensemble_q_function.py:90
        rew, term = _dev_f32(impl, rewards).reshape(-1), _dev_f32(impl, terminals).reshape(-1)
        assert rew.numel() == n and tgt.numel() == n and term.numel() == n
        g_rows, g = None, 0.0
        if isinstance(gamma, (torch.Tensor, np.ndarray)) and int(np.prod(tuple(gamma.shape))) > 1:
            g_rows = _dev_f32(impl, gamma).reshape(-1)
            assert g_rows.numel() == n
        else:
            g = float(gamma)
        out = torch.zeros(1, dtype=torch.float32, device=impl._device)
        impl.sync()   # the zero fill above ran on torch's stream
        impl._lib.td_error(q.data_ptr(), q.shape[1], rew.data_ptr(), tgt.data_ptr(), term.data_ptr(),
                           g_rows.data_ptr() if g_rows is not None else None, g, n, E, 1 if huber else 0, out.data_ptr(),
                           impl._stream)
        impl.sync()
        return out[0]


class EnsembleContinuousQFunction(_EnsembleBase):
    """ensemble_q_function.py:137-184 (continuous mean Q functions, mean_q_function.py:45-103)."""

    def _values(self, x, action) -> torch.Tensor:
        """Q_e(x, a) of every member of the view: [E, n] (a workspace: reduce / copy before the next call)."""
        self._check()
        impl = self._impl
        obs, act = _dev_f32(impl, x), _dev_f32(impl, action)
        assert obs.ndim == 2 and act.ndim == 2 and obs.shape[0] == act.shape[0]
        n, O, A = obs.shape[0], obs.shape[1], impl._action_size
        assert O == impl._observation_shape[0] and act.shape[1] == A
        impl.sync()
        rows = impl.ws("qf_x", n, O + A)
        impl._lib.concat_rows(obs.data_ptr(), O, act.data_ptr(), A, None, 0.0, 0.0, 0.0, rows.data_ptr(), O + A, n, 1, O, A,
                              impl._stream)
        _, q = impl._critic_rows_forward(self._which, rows, n, f"qf_{self._which}", members=self._members,
                                         member0=self._member0, train=False)
        return q

    def __call__(self, x, action, reduction: str = "mean") -> torch.Tensor:
        """forward (ensemble_q_function.py:141-151): (n, 1), or (E, n, 1) with reduction="none"."""
        return self._reduce(self._values(x, action), 1, reduction, 0.75)

    forward = __call__

    def compute_target(self, x, action, reduction: str = "min", lam: float = 0.75) -> torch.Tensor:
        """ensemble_q_function.py:153-184 for mean Q functions: the same values under another default reduction."""
        return self._reduce(self._values(x, action), 1, reduction, lam)

    def compute_error(self, observations, actions, rewards, target, terminals, gamma=0.99) -> torch.Tensor:
        """sum over members of mean_b (Q_e(s, a) - (r + gamma * target * (1 - terminal)))^2
        (ensemble_q_function.py:81-106, mean_q_function.py:74-87); `gamma` a float or a per-row tensor."""
        return self._td(self._values(observations, actions), rewards, target, terminals, gamma, huber=False)


class EnsembleDiscreteQFunction(_EnsembleBase):
    """ensemble_q_function.py:114-134 (discrete mean Q functions, mean_q_function.py:13-42)."""

    def _values(self, x) -> torch.Tensor:
        """Q_e(x, .) of every member: [E, n * A]."""
        self._check()
        from .algos.torch.dqn_impl import _EvalBatch

        impl = self._impl
        assert self._members == impl._n_critics and self._member0 == 0, "member views of discrete critics: use reduction='none'"
        xs = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
        db = impl.load_batch(_EvalBatch(xs))
        _, q = impl._forward(self._which, db, "obs", f"qf_{self._which}", False)
        return q.view(q.shape[0], -1)

    def __call__(self, x, reduction: str = "mean") -> torch.Tensor:
        """forward (ensemble_q_function.py:115-119): (n, A), or (E, n, A) with reduction="none"."""
        return self._reduce(self._values(x), self._impl._action_size, reduction, 0.75)

    forward = __call__

    def _picked(self, x, action) -> torch.Tensor:
        """pick_value_by_action (q_functions/utility.py:7-14) per member: [E, n]."""
        impl, A = self._impl, self._impl._action_size
        q = self._values(x)
        impl.sync()
        a = torch.as_tensor(np.asarray(action.detach().cpu() if isinstance(action, torch.Tensor) else action)).reshape(-1)
        a = a.to(impl._device, torch.int64)
        with torch.cuda.stream(impl._stream_obj):
            picked = q.view(q.shape[0], -1, A).gather(2, a.view(1, -1, 1).expand(q.shape[0], -1, 1)).squeeze(2).contiguous()
        return picked

    def compute_target(self, x, action=None, reduction: str = "min", lam: float = 0.75) -> torch.Tensor:
        """ensemble_q_function.py:108-134: all actions (n, A) when `action` is None, else the picked value (n, 1)."""
        if action is None:
            return self._reduce(self._values(x), self._impl._action_size, reduction, lam)
        return self._reduce(self._picked(x, action), 1, reduction, lam)

    def compute_error(self, observations, actions, rewards, target, terminals, gamma=0.99) -> torch.Tensor:
        """sum over members of mean_b huber(Q_e(s, a_b) - (r + gamma * target * (1 - terminal)))
        (ensemble_q_function.py:81-106, mean_q_function.py:26-42)."""
        return self._td(self._picked(observations, actions), rewards, target, terminals, gamma, huber=True)
