"""Actor-critic impl base: critic ensemble + policy over flat arenas; mirrors DDPGBaseImpl
(d3rlpy/algos/torch/ddpg_impl.py:29-229) — same hook names, hand-written backward instead of
autograd, one launch per layer for all ensemble members."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from ...nets import DenseNet
from ...q_functions import EnsembleContinuousQFunction, _ModuleView
from .base import ImplBase

# counter slots (device int32): 0 = noise epoch; the rest are Adam step counts
C_DRAW, C_CRITIC, C_ACTOR, C_TEMP, C_ALPHA, C_IMITATOR = 0, 1, 2, 3, 4, 5


class _OptimView:
    """`torch.optim.Adam.state_dict()` / `load_state_dict()`-shaped view of the fused optimizer state of one parameter
    group.  `sd_of(which)` returns the OrderedDict of named tensors for which in {params, exp_avg, exp_avg_sq};
    parameter index i of the torch optimizer == i-th key of the module's state_dict (registration order)."""

    def __init__(self, sd_of, step: torch.Tensor, lr: float):
        self._sd_of, self._step, self.lr = sd_of, step, lr

    def state_dict(self):
        keys = list(self._sd_of("params").keys())
        m, v = self._sd_of("exp_avg"), self._sd_of("exp_avg_sq")
        step = int(self._step.item())
        state = {i: {"step": torch.tensor(float(step)), "exp_avg": m[k].detach().cpu().clone(),
                     "exp_avg_sq": v[k].detach().cpu().clone()} for i, k in enumerate(keys)}
        # same group keys as torch 2.11's Adam (tests/golden/checkpoint_keys.json, recorded from the reference)
        group = {"lr": self.lr, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(keys)))}
        return {"state": state if step > 0 else {}, "param_groups": [group]}

    def load_state_dict(self, sd):
        keys = list(self._sd_of("params").keys())
        state = sd.get("state", {})
        m, v = self._sd_of("exp_avg"), self._sd_of("exp_avg_sq")
        step = 0
        with torch.no_grad():
            for i, k in enumerate(keys):
                st = state.get(i, state.get(str(i)))
                if st is None:
                    m[k].zero_()
                    v[k].zero_()
                    continue
                m[k].copy_(torch.as_tensor(st["exp_avg"]).to(m[k].device, torch.float32).reshape(m[k].shape))
                v[k].copy_(torch.as_tensor(st["exp_avg_sq"]).to(v[k].device, torch.float32).reshape(v[k].shape))
                step = int(float(st["step"]))
            self._step.fill_(step)


def _net_optim(net: DenseNet, lr: float) -> _OptimView:
    return _OptimView(lambda which: net.arena.state_dict(which), net.arena.step, lr)


class DDPGBaseImpl(ImplBase):
    def __init__(self, observation_shape, action_size, actor_learning_rate, critic_learning_rate,
                 actor_hidden: Sequence[int], critic_hidden: Sequence[int], gamma, tau, n_critics, use_gpu=0,
                 scaler=None, action_scaler=None, reward_scaler=None, seed: int = 0, precision: str = "fp32",
                 n_quantiles: int = 0, **kw):
        super().__init__(observation_shape, action_size, use_gpu, scaler, action_scaler, reward_scaler, **kw)
        self._precision = precision
        # > 0: every critic is a ContinuousQRQFunction (qr_q_function.py:91-165) — head Linear(feature, n_quantiles),
        # Q = mean over the quantiles; only the impls that set SUPPORTS_QR consume it
        self._n_quantiles = int(n_quantiles or 0)
        if self._n_quantiles and not getattr(self, "SUPPORTS_QR", False):
            raise ValueError(f"{type(self).__name__}: quantile-regression critics are not on the accelerated path")
        if self._n_quantiles > 32:
            raise ValueError("continuous quantile-regression critics support up to 32 quantiles (narrow-head kernels)")
        assert len(self._observation_shape) == 1, "vector observations only for actor-critic impls"
        self._actor_learning_rate, self._critic_learning_rate = actor_learning_rate, critic_learning_rate
        self._actor_hidden, self._critic_hidden = list(actor_hidden), list(critic_hidden)
        self._gamma, self._tau, self._n_critics = gamma, tau, n_critics
        self._gen = torch.Generator().manual_seed(seed)
        self._seed = seed
        self._q_func: Optional[DenseNet] = None
        self._policy: Optional[DenseNet] = None

    # ------------------------------------------------------------------ build
    def build(self) -> None:
        O, A = self._observation_shape[0], self._action_size
        self._q_func = DenseNet(O + A, self._critic_hidden, [("_fc", max(1, self._n_quantiles))], self._n_critics,
                                self._device,
                                trunk_prefix="_encoder.", member_key="_q_funcs.{e}.{name}", with_target=True,
                                seed_gen=self._gen, precision=self._precision)
        self._build_actor()
        self._q_func.arena.step = self._counters[C_CRITIC:C_CRITIC + 1]
        self._policy.arena.step = self._counters[C_ACTOR:C_ACTOR + 1]
        for net in (self._q_func, self._policy):
            net.refresh_shadow("params", self._stream)
            net.refresh_shadow("target", self._stream)
        self.sync()

    def _build_actor(self) -> None:
        raise NotImplementedError

    # ------------------------------------------------------------------ reference-visible properties
    @property
    def policy(self):
        return _ModuleView(self._policy)

    @property
    def targ_policy(self):
        return _ModuleView(self._policy, "target")

    @property
    def q_function(self):
        """Callable like the reference's EnsembleContinuousQFunction (`(x, action, reduction)`, `compute_error`,
        `compute_target`, `q_funcs`), see d3rlpy_b200/q_functions.py."""
        return EnsembleContinuousQFunction(self)

    @property
    def targ_q_function(self):
        return EnsembleContinuousQFunction(self, "target")

    @property
    def policy_optim(self):
        return _net_optim(self._policy, self._actor_learning_rate)

    @property
    def q_function_optim(self):
        return _net_optim(self._q_func, self._critic_learning_rate)

    # ------------------------------------------------------------------ checkpoints (algos/torch/base.py:137-142)
    def _checkpoint_views(self):
        """attribute name -> state_dict view, the keys torch_utility.get_state_dict (torch_utility.py:97-110) finds on
        the reference impl."""
        return {"_q_func": self.q_function, "_targ_q_func": self.targ_q_function, "_policy": self.policy,
                "_targ_policy": self.targ_policy, "_critic_optim": self.q_function_optim,
                "_actor_optim": self.policy_optim}

    # ------------------------------------------------------------------ shared program pieces
    def _critic_rows_forward(self, which: str, x, rows: int, tag: str, members=None, member0=0, train=True, stream=None):
        """Runs the critic trunk+head on `rows` shared input rows; returns (ctx, q[E,rows]) — (ctx, theta[E,rows,n])
        for quantile-regression critics."""
        E = members or self._n_critics
        ctx = self._q_func.ctx(tag, rows, E, train)
        q = self.ws(f"{tag}_q", E, rows, self._n_quantiles) if self._n_quantiles else self.ws(f"{tag}_q", E, rows)
        self._q_func.forward(which, x, self._q_func.in_dim, rows, ctx, q, self._stream if stream is None else stream,
                             member0=member0)
        return ctx, q

    def update_critic_target(self) -> None:
        """soft_sync(targ_q_func, q_func, tau) (ddpg_impl.py:201-204)."""
        a = self._q_func.arena
        self._lib.soft_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, self._tau, self._stream)
        self._q_func.refresh_shadow("target", self._stream)

    def update_actor_target(self) -> None:
        """soft_sync(targ_policy, policy, tau) (ddpg_impl.py:206-209)."""
        a = self._policy.arena
        self._lib.soft_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, self._tau, self._stream)
        self._policy.refresh_shadow("target", self._stream)

    def _tick(self, *slots):
        mask = 0
        for s in slots:
            mask |= 1 << s
        self._lib.tick(self._counters.data_ptr(), self.N_COUNTERS, mask, self._stream)

    def _metrics_dict(self, names):
        vals = self.read_slots_after_program()
        return {n: np.float32(vals[i]) for i, n in names}

    # ------------------------------------------------------------------ evaluation API (algos/torch/base.py:52-84,
    # algos/torch/utility.py:51-78): used by scorers / deployment, eager launches + one sync
    def _eval_obs(self, x) -> torch.Tensor:
        """numpy / tensor observations -> float32 device rows with the observation scaler applied."""
        t = torch.as_tensor(np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x), dtype=torch.float32)
        assert t.ndim > 1, "Input must have batch dimension."
        with torch.cuda.stream(self._stream_obj):
            d = t.to(self._device, non_blocking=False).contiguous()
        if self._vector_scaler() is not None:   # StandardScaler / MinMaxScaler.transform (torch_api scaler_targets=["x"])
            mean, std, eps = self._scaler_params()
            self._lib.standardize(d.data_ptr(), mean.data_ptr(), std.data_ptr(), eps, d.shape[0], d.shape[1],
                                  self._stream)
        return d

    def _policy_head(self, obs: torch.Tensor, head_tanh: bool) -> torch.Tensor:
        n = obs.shape[0]
        out = self.ws("eval_head", 1, n, self._policy.head_out)
        self._policy.forward("params", obs, obs.shape[1], n, self._policy.ctx("eval_pi", n, 1, False), out,
                             self._stream, head_tanh=head_tanh)
        return out

    def _predict_best_action(self, obs: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def predict_best_action(self, x, normalized: bool = False) -> np.ndarray:
        """TorchImplBase.predict_best_action (algos/torch/base.py:50-64): greedy action, mapped back to the original
        range by `action_scaler.reverse_transform` unless `normalized`."""
        obs = self._eval_obs(x)
        with torch.cuda.stream(self._stream_obj):
            a = self._predict_best_action(obs).contiguous()
        if not normalized:
            self.unscale_actions(a)
        self.sync()
        return a.detach().cpu().numpy()

    def predict_value(self, x, action, with_std: bool = False):
        """ContinuousQFunctionMixin.predict_value (algos/torch/utility.py:51-78): mean (and std) over members."""
        obs = self._eval_obs(x)
        with torch.cuda.stream(self._stream_obj):
            act = torch.as_tensor(np.asarray(action), dtype=torch.float32).to(self._device).contiguous()
        assert obs.shape[0] == act.shape[0]
        if self._action_scaler is not None:   # torch_api(action_scaler_targets=["action"]), algos/torch/utility.py:51-55
            mn, mx = self._action_scaler_params()
            self._lib.scale_actions(act.data_ptr(), mn.data_ptr(), mx.data_ptr(), act.shape[0], self._action_size,
                                    self._stream)
        n, O, A = obs.shape[0], obs.shape[1], self._action_size
        rows = self.ws("eval_x", n, O + A)
        self._lib.concat_rows(obs.data_ptr(), O, act.data_ptr(), A, None, 0.0, 0.0, 0.0, rows.data_ptr(), O + A, n, 1,
                              O, A, self._stream)
        _, q = self._critic_rows_forward("params", rows, n, "eval_q", train=False)
        self.sync()
        values = q.detach().cpu().numpy()                     # [E, n]
        if self._n_quantiles:
            values = values.mean(axis=2)                      # ContinuousQRQFunction.forward: mean over quantiles
        mean, std = values.mean(axis=0), values.std(axis=0)
        return (mean, std) if with_std else mean
