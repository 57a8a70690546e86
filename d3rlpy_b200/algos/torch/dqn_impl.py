"""DQN / DoubleDQN / DiscreteCQL on B200: mirrors DQNImpl, DoubleDQNImpl (d3rlpy/algos/torch/dqn_impl.py:19-171)
and DiscreteCQLImpl (d3rlpy/algos/torch/cql_impl.py:246-302).

One update = online-net forward on s' (Double-DQN action), target-net forward on s', ONE online-net
forward + backward on s shared by the Huber TD term and the conservative term (the reference evaluates
`self._q_func(obs_t)` twice, cql_impl.py:295,300, with identical values), Adam, and the hard target copy
when `grad_step % target_update_interval == 0` (pre-increment, dqn.py:130-131).

`n_quantiles > 0` switches every member to the quantile-regression Q function (QRQFunctionFactory,
models/q_functions.py:123-165; DiscreteQRQFunction, q_functions/qr_q_function.py:22-88): the head emits
A * n_quantiles values per sample (a dense layer on the GEMM kernels), Q(s, a) is the mean over the quantiles, the
target is the quantile vector of the member with the smallest mean, the TD term is the quantile Huber loss
(csrc/qr.cu)."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from ...nets import ConvNet, DenseNet
from .base import ImplBase
from .ddpg_impl import C_CRITIC, _ModuleView, _net_optim

M_LOSS = 0
S_LOSS = 0


class DQNImpl(ImplBase):
    POLICY_KIND = "discrete"
    DISCRETE = True
    DOUBLE = False
    CONSERVATIVE = False

    def __init__(self, observation_shape, action_size, learning_rate, hidden: Sequence[int], gamma, n_critics,
                 feature_size: int = 512, filters=None, use_gpu=0, scaler=None, reward_scaler=None, seed: int = 0,
                 precision: str = "fp32", alpha: float = 1.0, n_quantiles: int = 0, **kw):
        super().__init__(observation_shape, action_size, use_gpu, scaler, None, reward_scaler, **kw)
        self._n_quantiles = int(n_quantiles or 0)
        self._learning_rate, self._hidden, self._gamma, self._n_critics = learning_rate, list(hidden), gamma, n_critics
        self._feature_size, self._filters = feature_size, filters
        self._precision, self._alpha = precision, alpha
        self._gen = torch.Generator().manual_seed(seed)
        self._seed = seed
        self._q_func = None

    @property
    def _pixel(self) -> bool:
        return len(self._observation_shape) == 3

    def build(self) -> None:
        A, E = self._action_size * max(1, self._n_quantiles), self._n_critics
        if self._pixel:
            sc = self._scaler
            is_pixel_scaler = sc == "pixel" or getattr(sc, "TYPE", None) == "pixel"
            if sc is not None and not is_pixel_scaler:
                raise ValueError("pixel observations support scaler=None or 'pixel'")
            self._q_func = ConvNet(self._observation_shape, [("_fc", A)], E, self._device,
                                   feature_size=self._feature_size, filters=self._filters,
                                   member_key="_q_funcs.{e}.{name}", with_target=True, seed_gen=self._gen,
                                   precision=self._precision, input_divisor=255.0 if is_pixel_scaler else 1.0)
        else:
            self._q_func = DenseNet(self._observation_shape[0], self._hidden, [("_fc", A)], E, self._device,
                                    trunk_prefix="_encoder.", member_key="_q_funcs.{e}.{name}", with_target=True,
                                    seed_gen=self._gen, precision=self._precision)
        self._q_func.arena.step = self._counters[C_CRITIC:C_CRITIC + 1]
        self._q_func.refresh_shadow("params", self._stream)
        self._q_func.refresh_shadow("target", self._stream)
        self.sync()

    # ------------------------------------------------------------------ reference-visible properties
    @property
    def q_function(self):
        from ...q_functions import EnsembleDiscreteQFunction

        return EnsembleDiscreteQFunction(self)

    @property
    def targ_q_function(self):
        from ...q_functions import EnsembleDiscreteQFunction

        return EnsembleDiscreteQFunction(self, "target")

    @property
    def q_function_optim(self):
        return _net_optim(self._q_func, self._learning_rate)

    def _checkpoint_views(self):
        return {"_q_func": self.q_function, "_targ_q_func": self.targ_q_function, "_optim": self.q_function_optim}

    # ------------------------------------------------------------------ program pieces
    def _forward(self, which, db, field, tag, train, stream=None):
        """Q values [E, B, A] (quantiles [E, B, A * n_quantiles]) of the chosen parameter set on obs / next_obs."""
        B, A, E = db.B, self._action_size * max(1, self._n_quantiles), self._n_critics
        st = self._stream if stream is None else stream
        q = self.ws(f"{tag}_q", E, B, A)
        if self._pixel:
            ctx = self._q_func.ctx(tag, B, E, train)
            self._q_func.forward(which, db.ptr(field), B, ctx, q, st)
        else:
            ctx = self._q_func.ctx(tag, B, E, train)
            self._q_func.forward(which, db.ptr(field), db.O, B, ctx, q, st)
        return ctx, q

    def _side_streams(self):
        if getattr(self, "_side_objs", None) is None:
            self._side_objs = (torch.cuda.Stream(device=self._device), torch.cuda.Stream(device=self._device))
        return self._side_objs[0].cuda_stream, self._side_objs[1].cuda_stream

    def _p_next_values(self, db):
        """The target network on s' and (Double DQN) the online network on s', each on its own graph branch; the caller
        joins them with `_join_branches` before `_p_target_kernel`.  Together with the online network on s (main
        stream, `_p_loss`) these are three independent chains of small launches."""
        L, st = self._lib, self._stream
        s1, s2 = self._side_streams()
        L.stream_fork(st, s1)
        _, q_t = self._forward("target", db, "next_obs", "tq", False, stream=s1)
        self._pending_joins = [s1]
        q_sel = q_t
        if self.DOUBLE:
            L.stream_fork(st, s2)
            q_sel = self._forward("params", db, "next_obs", "oq", False, stream=s2)[1]
            self._pending_joins.append(s2)
        return q_sel, q_t

    def _join_branches(self):
        for s in getattr(self, "_pending_joins", []):
            self._lib.stream_join(self._stream, s)
        self._pending_joins = []

    def _p_target_kernel(self, db, q_sel, q_t):
        B, A, E, L, st = db.B, self._action_size, self._n_critics, self._lib, self._stream
        nq = self._n_quantiles
        if nq:
            q_tpn = self.ws("q_tpn", B, nq)
            L.qr_target(q_sel.data_ptr(), B * A * nq, q_t.data_ptr(), B * A * nq, q_tpn.data_ptr(), B, A, nq, E, st)
            return q_tpn
        q_tpn = self.ws("q_tpn", B)
        L.dqn_target(q_sel.data_ptr(), B * A, q_t.data_ptr(), B * A, q_tpn.data_ptr(), B, A, E, st)
        return q_tpn

    def _p_target(self, db):
        """DQNImpl.compute_target (dqn_impl.py:133-141) / DoubleDQNImpl.compute_target (:162-171)."""
        q_sel, q_t = self._p_next_values(db)
        self._join_branches()
        return self._p_target_kernel(db, q_sel, q_t)

    def _p_update(self, db):
        """One whole update: the three forward passes side by side, then target, loss, backward, Adam."""
        q_sel, q_t = self._p_next_values(db)
        fwd = self._forward("params", db, "obs", "lq", True)
        self._join_branches()
        self._p_loss(db, self._p_target_kernel(db, q_sel, q_t), fwd=fwd)

    def _p_loss(self, db, q_tpn, step=True, fwd=None):
        """compute_loss (dqn_impl.py:113-131; cql_impl.py:279-302) [+ backward + Adam (dqn_impl.py:97-111)]."""
        B, A, E, L, st = db.B, self._action_size, self._n_critics, self._lib, self._stream
        ctx, q = fwd if fwd is not None else self._forward("params", db, "obs", "lq", True)
        nq = self._n_quantiles
        dq = self.ws("dq", E, B, A * max(1, nq))
        inv_b = 1.0 / (B * self.world_size)
        cons = 1 if self.CONSERVATIVE else 0
        if nq:
            L.qr_loss(q.data_ptr(), B * A * nq, q_tpn.data_ptr(), db.ptr("act"), db.ptr("rew"), db.ptr("term"),
                      db.ptr("nsteps"), self._gamma, self._alpha, dq.data_ptr(), B * A * nq,
                      self.ws("qr_partials", 2 * B).data_ptr(), self.sums_ptr(S_LOSS), B, A, nq, E, inv_b, cons, st)
        else:
            L.dcql_loss(q.data_ptr(), B * A, q_tpn.data_ptr(), db.ptr("act"), db.ptr("rew"), db.ptr("term"),
                        db.ptr("nsteps"), self._gamma, self._alpha, dq.data_ptr(), B * A, self.sums_ptr(S_LOSS), B, A,
                        E, inv_b, cons, st)
        self._allreduce(self._slots[32 + S_LOSS:32 + S_LOSS + 2])
        L.dcql_finalize(self.sums_ptr(S_LOSS), inv_b, self._alpha, cons, self.metric_ptr(M_LOSS), st)
        if step:
            if self._pixel:
                self._q_func.backward(B, ctx, dq, st)
            else:
                self._q_func.backward(db.ptr("obs"), db.O, B, ctx, dq, st)
            self._allreduce(self._q_func.arena.grads)
            self._q_func.adam(self._learning_rate, st)

    def _p_hard_sync(self):
        a = self._q_func.arena
        self._lib.hard_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, self._stream)
        self._q_func.refresh_shadow("target", self._stream)

    def _allreduce(self, t):
        if self.world_size > 1:
            from ...parallel import allreduce_sum

            allreduce_sum(t, self._stream_obj)

    def _tick(self, *slots):
        mask = 0
        for s in slots:
            mask |= 1 << s
        self._lib.tick(self._counters.data_ptr(), self.N_COUNTERS, mask, self._stream)

    # ------------------------------------------------------------------ fused update (DQN._update, dqn.py:127-132)
    def update_fused(self, batch, sync_target: bool):
        self.update_fused_async(batch, sync_target)
        return {"loss": np.float32(self.read_slots_after_program()[M_LOSS])}

    def update_fused_async(self, batch, sync_target: bool):
        db = self.load_batch(batch, defer=True)

        def program():
            self._tick(C_CRITIC)
            self.zero_slots()
            self._p_update(db)
            if sync_target:
                self._p_hard_sync()

        self.run_program(("dqn", db.B, sync_target), program)
        return [(M_LOSS, "loss")]

    # ------------------------------------------------------------------ reference hooks (eager)
    def update(self, batch) -> np.ndarray:
        db = self.load_batch(batch)
        self._tick(C_CRITIC)
        self.zero_slots()
        self._p_update(db)
        return self.read_slots()[M_LOSS].copy()

    def compute_target(self, batch) -> torch.Tensor:
        db = self.load_batch(batch)
        q = self._p_target(db)
        self.sync()
        return q.view(-1, self._n_quantiles or 1).clone()  # (B, 1), or (B, n_quantiles) like qr_q_function.py:80-88

    def compute_loss(self, batch, q_tpn: torch.Tensor) -> torch.Tensor:
        db = self.load_batch(batch)
        self.zero_slots()
        self._p_loss(db, q_tpn.to(self._device).reshape(-1).contiguous(), step=False)
        self.sync()
        return self._slots[M_LOSS].clone()

    def update_target(self) -> None:
        """hard_sync(targ_q_func, q_func) (dqn_impl.py:143-145)."""
        self._p_hard_sync()


class _EvalBatch:
    """Staging shim: observations only (the evaluation API reuses the minibatch upload path)."""

    def __init__(self, x, n_act=1):
        x = np.asarray(x)
        n = x.shape[0]
        self.observations = self.next_observations = x
        self.actions = np.zeros(n, np.float32)
        self.rewards = self.terminals = np.zeros((n, 1), np.float32)
        self.n_steps = np.ones((n, 1), np.float32)


def _dqn_q_values(impl, x) -> np.ndarray:
    """Q(s, .) of every member: [E, n, A] (eager forward of the online network, one sync)."""
    db = impl.load_batch(_EvalBatch(x))
    _, q = impl._forward("params", db, "obs", "eval", False)
    nq = impl._n_quantiles
    if nq:  # DiscreteQRQFunction.forward: mean over the quantiles (qr_q_function.py:44-48)
        E, n, A = q.shape[0], q.shape[1], impl._action_size
        v = impl.ws("eval_values", E, n, A)
        impl._lib.qr_values(q.data_ptr(), n * A * nq, v.data_ptr(), n * A, n, A, nq, E, impl._stream)
        q = v
    impl.sync()
    return q.detach().cpu().numpy()


def _dqn_predict_best_action(self, x) -> np.ndarray:
    """argmax_a mean_e Q_e(s, a) (dqn_impl.py:147-149)."""
    return _dqn_q_values(self, x).mean(axis=0).argmax(axis=1)


def _dqn_predict_value(self, x, action, with_std: bool = False):
    """DiscreteQFunctionMixin.predict_value (algos/torch/utility.py:22-49)."""
    values = _dqn_q_values(self, x)                                # [E, n, A]
    a = np.asarray(action).reshape(-1).astype(np.int64)
    picked = values[:, np.arange(values.shape[1]), a]              # [E, n]
    mean, std = picked.mean(axis=0), picked.std(axis=0)
    return (mean, std) if with_std else mean


DQNImpl.predict_best_action = _dqn_predict_best_action
DQNImpl.sample_action = _dqn_predict_best_action
DQNImpl.predict_value = _dqn_predict_value


class DoubleDQNImpl(DQNImpl):
    DOUBLE = True


class DiscreteCQLImpl(DoubleDQNImpl):
    CONSERVATIVE = True

    def _compute_conservative_loss(self, obs_t, act_t) -> torch.Tensor:
        """cql_impl.py:290-302: mean_b(logsumexp_a Qbar(s) - Qbar(s, a_data)), Qbar = ensemble mean."""
        from types import SimpleNamespace

        n = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
        B = n(obs_t).shape[0]
        batch = SimpleNamespace(observations=n(obs_t), actions=n(act_t), next_observations=n(obs_t),
                                rewards=np.zeros((B, 1), np.float32), terminals=np.zeros((B, 1), np.float32),
                                n_steps=np.ones((B, 1), np.float32))
        db = self.load_batch(batch)
        self.zero_slots()
        self._p_loss(db, self.ws("zero_tpn", B * max(1, self._n_quantiles)), step=False)
        self.sync()
        return (self._slots[32 + S_LOSS + 1] / B).clone()
