"""IQL on B200: mirrors IQLImpl (d3rlpy/algos/torch/iql_impl.py:20-200) over the same dense-network kernels as the
other actor-critic impls — critic ensemble + value function under ONE Adam (`_build_critic_optim`, :92-99), a
non-squashed Gaussian policy with a learnable logstd parameter (`create_non_squashed_normal_policy(min_logstd=-5,
max_logstd=2, use_std_parameter=True)`, :74-82).  One update (`IQL._update`, algos/iql.py:186-199):

  critic step   Q TD loss against V(s') (no target network on V, :101-117) + expectile value loss against
                min_e Q'_e(s, a) (:130-141), one backward, one Adam step over both networks
  actor step    -mean(w log pi(a|s)), w = clamp(exp(weight_temp (min Q' - V)), max_weight) with the UPDATED V (:109-128)
  target        soft_sync(targ_q, q) only

No random draws in the update.  Loss tails: csrc/iql.cu."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from ...arena import ParamArena
from ...nets import DenseNet
from .ddpg_impl import C_ACTOR, C_CRITIC, DDPGBaseImpl, _ModuleView, _OptimView

M_CRITIC, M_VALUE, M_ACTOR = 0, 1, 2
S_TD = 0
MIN_LOGSTD, MAX_LOGSTD = -5.0, 2.0


class _PolicyView:
    """`state_dict()` of NonSquashedNormalPolicy: the `_logstd` nn.Parameter is registered on the module itself, so it
    comes BEFORE the encoder and `_mu` entries (it is also parameter 0 of the actor optimizer)."""

    def __init__(self, impl: "IQLImpl", which: str = "params"):
        self._impl, self._which = impl, which

    def state_dict(self):
        sd = OrderedDict(self._impl._logstd.state_dict(self._which))
        sd.update(self._impl._policy.arena.state_dict(self._which))
        return sd

    def load_state_dict(self, sd):
        self._impl._logstd.load_state_dict({"_logstd": sd["_logstd"]}, self._which)
        _ModuleView(self._impl._policy, self._which).load_state_dict(sd)

    def parameters(self):
        return list(self.state_dict().values())


class IQLImpl(DDPGBaseImpl):
    POLICY_KIND = "normal"

    def __init__(self, *, value_hidden, expectile=0.7, weight_temp=3.0, max_weight=100.0, **kw):
        super().__init__(**kw)
        self._value_hidden = list(value_hidden)
        self._expectile, self._weight_temp, self._max_weight = expectile, weight_temp, max_weight
        self._value = None
        self._logstd = None

    def build(self) -> None:
        if self.world_size > 1:
            raise NotImplementedError("IQL: data-parallel exchange is not wired (single GPU only)")
        super().build()
        O = self._observation_shape[0]
        self._value = DenseNet(O, self._value_hidden, [("_fc", 1)], 1, self._device, trunk_prefix="_encoder.",
                               seed_gen=self._gen, precision=self._precision)
        self._value.arena.step = self._counters[C_CRITIC:C_CRITIC + 1]   # one Adam over critics + value function
        self._value.refresh_shadow("params", self._stream)
        self.sync()

    def _build_actor(self) -> None:
        O, A = self._observation_shape[0], self._action_size
        self._policy = DenseNet(O, self._actor_hidden, [("_mu", A)], 1, self._device, trunk_prefix="_encoder.",
                                with_target=True, seed_gen=self._gen, precision=self._precision)
        self._logstd = ParamArena([("_logstd", (1, A))], 1, self._device, with_target=True)   # zeros (policies.py:56-57)
        self._logstd.step = self._counters[C_ACTOR:C_ACTOR + 1]

    # ------------------------------------------------------------------ reference-visible properties
    @property
    def policy(self):
        return _PolicyView(self)

    @property
    def targ_policy(self):
        return _PolicyView(self, "target")

    @property
    def value_function(self):
        return _ModuleView(self._value)

    @property
    def policy_optim(self):
        def sd_of(which):
            sd = OrderedDict(self._logstd.state_dict(which))
            sd.update(self._policy.arena.state_dict(which))
            return sd

        return _OptimView(sd_of, self._policy.arena.step, self._actor_learning_rate)

    @property
    def q_function_optim(self):
        def sd_of(which):   # parameter order of `q_func_params + v_func_params` (iql_impl.py:95-99)
            sd = OrderedDict(self._q_func.arena.state_dict(which))
            sd.update(("_value_func." + k, v) for k, v in self._value.arena.state_dict(which).items())
            return sd

        return _OptimView(sd_of, self._q_func.arena.step, self._critic_learning_rate)

    def _checkpoint_views(self):
        views = super()._checkpoint_views()
        views["_value_func"] = self.value_function
        return views

    # ------------------------------------------------------------------ program pieces
    def _rows(self, db):
        B, O, A = db.B, db.O, self._action_size
        xc = self.ws("xc", B, O + A)
        self._lib.concat_rows(db.ptr("obs"), O, db.ptr("act"), A, None, 0.0, 0.0, 0.0, xc.data_ptr(), O + A, B, 1, O, A,
                              self._stream)
        return xc

    def _v(self, db, field: str, tag: str, train: bool):
        B = db.B
        ctx = self._value.ctx(tag, B, 1, train)
        v = self.ws(f"{tag}_v", 1, B, 1)
        self._value.forward("params", db.ptr(field), db.O, B, ctx, v, self._stream)
        return ctx, v

    def _targ_q(self, xc, B):
        return self._critic_rows_forward("target", xc, B, "tq", train=False)[1]

    def _p_q_loss(self, db, xc, q_tpn):
        """compute_critic_loss (iql_impl.py:101-112): sum over members of batch-mean MSE against r + gamma^n V(s')."""
        B, A, E, L, st = db.B, self._action_size, self._n_critics, self._lib, self._stream
        acts, q = self._critic_rows_forward("params", xc, B, "cq")
        dq = self.ws("dq", E, B)
        inv_b = 1.0 / B
        L.critic_loss(q.data_ptr(), B, None, B, E, q_tpn.data_ptr(), db.ptr("rew"), db.ptr("term"), db.ptr("nsteps"),
                      self._gamma, None, None, 0, A, None, 0.0, dq.data_ptr(), B, self.sums_ptr(S_TD), None, B, E, inv_b,
                      1, st)
        L.cql_finalize(self.sums_ptr(S_TD), None, inv_b, E, 0.0, 0.0, 0, 0, self.metric_ptr(M_CRITIC), None, st)
        return acts, dq

    def _p_value_loss(self, db, q_t):
        """compute_value_loss (iql_impl.py:130-141)."""
        B = db.B
        ctx_v, v = self._v(db, "obs", "v", True)
        dv = self.ws("dv", 1, B, 1)
        self._lib.iql_value_loss(q_t.data_ptr(), B, self._n_critics, v.data_ptr(), self._expectile, 1.0 / B,
                                 dv.data_ptr(), self.metric_ptr(M_VALUE), B, self._stream)
        return ctx_v, dv

    def _p_critic(self, db, xc, q_t):
        B, O, A, st = db.B, db.O, self._action_size, self._stream
        _, v_next = self._v(db, "next_obs", "vn", False)                  # compute_target (iql_impl.py:114-117)
        acts, dq = self._p_q_loss(db, xc, v_next)
        ctx_v, dv = self._p_value_loss(db, q_t)
        self._q_func.backward(xc, O + A, B, acts, dq, st)
        self._q_func.adam(self._critic_learning_rate, st)
        self._value.backward(db.ptr("obs"), O, B, ctx_v, dv, st)
        self._value.adam(self._critic_learning_rate, st)

    def _p_actor_loss(self, db, q_t):
        """compute_actor_loss (iql_impl.py:109-128) -> (policy ctx, dL/dmu); accumulates dL/d_logstd."""
        B, O, A, st = db.B, db.O, self._action_size, self._stream
        _, v = self._v(db, "obs", "va", False)                            # the value function as just updated
        ctx_p = self._policy.ctx("pi", B, 1, True)
        mu = self.ws("pi_mu", 1, B, A)
        self._policy.forward("params", db.ptr("obs"), O, B, ctx_p, mu, st)
        dmu = self.ws("pi_dmu", 1, B, A)
        self._lib.iql_actor_loss(mu.data_ptr(), A, self._logstd.params.data_ptr(), db.ptr("act"), A, q_t.data_ptr(), B,
                                 self._n_critics, v.data_ptr(), self._weight_temp, self._max_weight, MIN_LOGSTD,
                                 MAX_LOGSTD, 1.0 / B, dmu.data_ptr(), A, self._logstd.grads.data_ptr(),
                                 self.metric_ptr(M_ACTOR), B, A, st)
        return ctx_p, dmu

    def _p_actor(self, db, q_t):
        B, O, st = db.B, db.O, self._stream
        ctx_p, dmu = self._p_actor_loss(db, q_t)
        self._policy.backward(db.ptr("obs"), O, B, ctx_p, dmu, st)
        self._policy.adam(self._actor_learning_rate, st)
        a = self._logstd
        self._lib.adam_step(a.params.data_ptr(), a.grads.data_ptr(), a.exp_avg.data_ptr(), a.exp_avg_sq.data_ptr(),
                            None, a.size, a.step.data_ptr(), self._actor_learning_rate, 0.9, 0.999, 1e-8, 0.0, 1, st)

    # ------------------------------------------------------------------ fused update (IQL._update)
    def update_fused(self, batch):
        return self._metrics_dict(self.update_fused_async(batch))

    def update_fused_async(self, batch):
        db = self.load_batch(batch, defer=True)

        def program():
            self._tick(C_CRITIC, C_ACTOR)
            self.zero_slots()
            xc = self._rows(db)
            q_t = self._targ_q(xc, db.B)     # min_e Q'_e(s, a): shared by the value loss and the actor weights
            self._p_critic(db, xc, q_t)
            self._p_actor(db, q_t)
            self.update_critic_target()

        self.run_program(("iql", db.B), program)
        return [(M_CRITIC, "critic_loss"), (M_VALUE, "value_loss"), (M_ACTOR, "actor_loss")]

    # ------------------------------------------------------------------ reference hooks (eager)
    def compute_target(self, batch) -> torch.Tensor:
        db = self.load_batch(batch)
        v = self._v(db, "next_obs", "vn", False)[1]
        self.sync()
        return v.view(-1, 1).clone()

    def compute_critic_loss(self, batch, q_tpn: torch.Tensor) -> torch.Tensor:
        db = self.load_batch(batch)
        self.zero_slots()
        self._p_q_loss(db, self._rows(db), q_tpn.to(self._device).reshape(-1).contiguous())
        self.sync()
        return self._slots[M_CRITIC].clone()

    def compute_value_loss(self, batch) -> torch.Tensor:
        db = self.load_batch(batch)
        self.zero_slots()
        self._p_value_loss(db, self._targ_q(self._rows(db), db.B))
        self.sync()
        return self._slots[M_VALUE].clone()

    def compute_actor_loss(self, batch) -> torch.Tensor:
        db = self.load_batch(batch)
        self.zero_slots()
        self._p_actor_loss(db, self._targ_q(self._rows(db), db.B))
        self.sync()
        self._logstd.grads.zero_()   # the hook only reports the loss; gradients belong to update_actor
        return self._slots[M_ACTOR].clone()

    def update_critic(self, batch):
        db = self.load_batch(batch)
        self._tick(C_CRITIC)
        self.zero_slots()
        xc = self._rows(db)
        self._p_critic(db, xc, self._targ_q(xc, db.B))
        vals = self.read_slots()
        return vals[M_CRITIC].copy(), vals[M_VALUE].copy()

    def update_actor(self, batch) -> np.ndarray:
        db = self.load_batch(batch)
        self._tick(C_ACTOR)
        self.zero_slots()
        self._p_actor(db, self._targ_q(self._rows(db), db.B))
        return self.read_slots()[M_ACTOR].copy()

    # ---- evaluation API
    def _predict_best_action(self, obs: torch.Tensor) -> torch.Tensor:
        """GaussianDistribution.mean = tanh(mu) (policies.py:176-181, distributions.py:83-84)."""
        return self._policy_head(obs, head_tanh=True)[0]

    def sample_action(self, x) -> np.ndarray:
        """dist.sample(): Normal(tanh(mu), exp(logstd)).rsample().clamp(-1, 1) (distributions.py:52-53); the draw is
        taken on the host from numpy's global stream."""
        mean = self.predict_best_action(x, normalized=True)
        p = self._logstd.params[:self._action_size].detach().cpu().numpy()
        std = np.exp(MIN_LOGSTD + (MAX_LOGSTD - MIN_LOGSTD) / (1.0 + np.exp(-p)))
        act = np.clip(mean + std * np.random.randn(*mean.shape), -1.0, 1.0).astype(np.float32)
        if self._action_scaler is None:
            return act
        with torch.cuda.stream(self._stream_obj):   # action_scaler.reverse_transform (algos/torch/base.py:77-79)
            d = torch.from_numpy(act).to(self._device).contiguous()
        self.unscale_actions(d)
        self.sync()
        return d.cpu().numpy()
