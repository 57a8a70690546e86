"""AWAC and CRR on B200: advantage-weighted actor steps over a non-squashed Gaussian policy, on the same dense-network
kernels and critic step as the other actor-critic impls.  Mirrors

  AWACImpl  d3rlpy/algos/torch/awac_impl.py:18-154 (SACImpl with a frozen temperature exp(log 1e-20), whose entropy
            term vanishes in float32; NonSquashedNormalPolicy with a logstd PARAMETER squashed into [-6, 0]; actor Adam
            with weight_decay 1e-4, algos/awac.py:105)
  CRRImpl   d3rlpy/algos/torch/crr_impl.py:17-191 (target action from the TARGET policy; NonSquashedNormalPolicy with a
            logstd HEAD clamped to [-20, 2]; binary / clipped-exponential weights; hard or soft target updates)

One update = critic step (target rows sampled from the policy, TD loss, backward, Adam) + actor step (ONE critic forward
over [data rows | n sampled rows per observation], advantage weights, weighted Gaussian log-likelihood, backward, Adam)
+ target syncs, captured as one CUDA graph.  Loss tails: csrc/awr.cu."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from ...arena import ParamArena
from ...nets import DenseNet
from .ddpg_impl import C_ACTOR, C_CRITIC, C_DRAW, DDPGBaseImpl, _ModuleView, _OptimView
from .iql_impl import _PolicyView

M_CRITIC, M_ACTOR, M_STD = 0, 1, 2
S_TD = 0


class _GaussActorImpl(DDPGBaseImpl):
    POLICY_KIND = "normal"
    LOGSTD_PARAM = False
    MIN_LOGSTD, MAX_LOGSTD = -20.0, 2.0
    TARGET_FROM_TARGET_POLICY = False
    MEMBER_REDUCE = 0        # 0 = min over members (AWAC), 1 = mean (CRR)
    ACTOR_WEIGHT_DECAY = 0.0

    def __init__(self, *, n_action_samples: int, **kw):
        super().__init__(**kw)
        self._n_action_samples = int(n_action_samples)
        self._logstd = None

    def build(self) -> None:
        if self.world_size > 1:
            raise NotImplementedError(f"{type(self).__name__}: data-parallel exchange is not wired (single GPU only)")
        super().build()

    def _build_actor(self) -> None:
        O, A = self._observation_shape[0], self._action_size
        heads = [("_mu", A)] if self.LOGSTD_PARAM else [("_mu", A), ("_logstd", A)]
        self._policy = DenseNet(O, self._actor_hidden, heads, 1, self._device, trunk_prefix="_encoder.",
                                with_target=True, seed_gen=self._gen, precision=self._precision)
        if self.LOGSTD_PARAM:
            self._logstd = ParamArena([("_logstd", (1, A))], 1, self._device, with_target=True)   # zeros (policies.py:56-57)
            self._logstd.step = self._counters[C_ACTOR:C_ACTOR + 1]

    # ------------------------------------------------------------------ reference-visible properties
    @property
    def policy(self):
        return _PolicyView(self) if self.LOGSTD_PARAM else _ModuleView(self._policy)

    @property
    def targ_policy(self):
        return _PolicyView(self, "target") if self.LOGSTD_PARAM else _ModuleView(self._policy, "target")

    @property
    def policy_optim(self):
        if not self.LOGSTD_PARAM:
            return super().policy_optim

        def sd_of(which):
            sd = OrderedDict(self._logstd.state_dict(which))
            sd.update(self._policy.arena.state_dict(which))
            return sd

        return _OptimView(sd_of, self._policy.arena.step, self._actor_learning_rate)

    def noise_layout(self, B):
        A, n = self._action_size, self._n_action_samples
        return {"target": ("normal", (B, A)), "weights": ("normal", (n, B, A))}

    # ------------------------------------------------------------------ program pieces
    def _logstd_ptr(self, which: str):
        if not self.LOGSTD_PARAM:
            return None
        return (self._logstd.params if which == "params" else self._logstd.target).data_ptr()

    def _head(self, which: str, db, field: str, tag: str, train: bool):
        B = db.B
        ctx = self._policy.ctx(tag, B, 1, train)
        head = self.ws(f"{tag}_head", 1, B, self._policy.head_out)
        self._policy.forward(which, db.ptr(field), db.O, B, ctx, head, self._stream)
        return ctx, head

    def _sample_rows(self, head, which: str, eps, obs_ptr, db, n: int, x_ptr: int):
        """rows [obs_b | clamp(tanh(mu) + std * eps, -1, 1)] (distributions.py:52-53), n per observation."""
        O, A = db.O, self._action_size
        self._lib.gauss_policy_rows(head.data_ptr(), self._policy.head_out, self._logstd_ptr(which), self.MIN_LOGSTD,
                                    self.MAX_LOGSTD, eps.data_ptr(), obs_ptr, O, x_ptr, O + A, db.B, n, O, A, self._stream)

    def _p_critic(self, db, backward=True, q_tpn=None):
        """compute_target + compute_critic_loss [+ backward + Adam] (sac_impl.py:148-162 / crr_impl.py:143-153,
        ddpg_impl.py:138-152)."""
        B, O, A, E, L, st = db.B, db.O, self._action_size, self._n_critics, self._lib, self._stream
        q_t = None
        if q_tpn is None:
            which = "target" if self.TARGET_FROM_TARGET_POLICY else "params"
            _, head = self._head(which, db, "next_obs", "tp", False)
            xt = self.ws("xt", B, O + A)
            self._sample_rows(head, which, self.noise_view("target", B), db.ptr("next_obs"), db, 1, xt.data_ptr())
            _, q_t = self._critic_rows_forward("target", xt, B, "tq", train=False)
        xc = self.ws("xc", B, O + A)
        L.concat_rows(db.ptr("obs"), O, db.ptr("act"), A, None, 0.0, 0.0, 0.0, xc.data_ptr(), O + A, B, 1, O, A, st)
        acts, q = self._critic_rows_forward("params", xc, B, "cq")
        dq = self.ws("dq", E, B)
        inv_b = 1.0 / B
        L.critic_loss(q.data_ptr(), B, q_t.data_ptr() if q_t is not None else None, B, E,
                      q_tpn.data_ptr() if q_tpn is not None else None, db.ptr("rew"), db.ptr("term"), db.ptr("nsteps"),
                      self._gamma, None, None, 0, A, None, 0.0, dq.data_ptr(), B, self.sums_ptr(S_TD), None, B, E, inv_b,
                      1, st)
        L.cql_finalize(self.sums_ptr(S_TD), None, inv_b, E, 0.0, 0.0, 0, 0, self.metric_ptr(M_CRITIC), None, st)
        if backward:
            self._q_func.backward(xc, O + A, B, acts, dq, st)
            self._q_func.adam(self._critic_learning_rate, st)

    def _weights_args(self):
        raise NotImplementedError

    def _loss_scale(self, B: int) -> float:
        raise NotImplementedError

    def _p_actor(self, db, step=True):
        """_compute_weights + compute_actor_loss [+ backward + Adam]."""
        B, O, A, E, n, L, st = db.B, db.O, self._action_size, self._n_critics, self._n_action_samples, self._lib, self._stream
        ctx_p, head = self._head("params", db, "obs", "pi", True)
        # ONE critic forward over [data rows | sampled rows]
        R = B * (1 + n)
        x = self.ws("xw", R, O + A)
        L.concat_rows(db.ptr("obs"), O, db.ptr("act"), A, None, 0.0, 0.0, 0.0, x.data_ptr(), O + A, B, 1, O, A, st)
        self._sample_rows(head, "params", self.noise_view("weights", B), db.ptr("obs"), db, n,
                          x.data_ptr() + 4 * (O + A) * B)
        _, q = self._critic_rows_forward("params", x, R, "wq", train=False)
        w = self.ws("aw", B)
        vr, mode, temp, max_w = self._weights_args()
        L.awr_weights(q.data_ptr(), R, q.data_ptr() + 4 * B, R, E, B, n, self.MEMBER_REDUCE, vr, mode, temp, max_w,
                      w.data_ptr(), st)
        H = self._policy.head_out
        d_head = self.ws("pi_dhead", 1, B, H)
        lp = self._logstd
        L.gauss_wll_loss(head.data_ptr(), H, self._logstd_ptr("params"), db.ptr("act"), A, w.data_ptr(), self.MIN_LOGSTD,
                         self.MAX_LOGSTD, self._loss_scale(B), d_head.data_ptr(), H,
                         lp.grads.data_ptr() if lp is not None else None, self.metric_ptr(M_ACTOR),
                         self.metric_ptr(M_STD) if lp is not None else None, B, A, st)
        if not step:
            return
        self._policy.backward(db.ptr("obs"), O, B, ctx_p, d_head, st)
        wd = self.ACTOR_WEIGHT_DECAY
        if wd == 0.0:
            self._policy.adam(self._actor_learning_rate, st)
        else:
            a = self._policy.arena
            L.adam_step_wd(a.params.data_ptr(), a.grads.data_ptr(), a.exp_avg.data_ptr(), a.exp_avg_sq.data_ptr(), None,
                           a.size, a.step.data_ptr(), self._actor_learning_rate, 0.9, 0.999, 1e-8, wd, 0.0, 1, st)
            self._policy.refresh_shadow("params", st)
        if lp is not None:
            L.adam_step_wd(lp.params.data_ptr(), lp.grads.data_ptr(), lp.exp_avg.data_ptr(), lp.exp_avg_sq.data_ptr(),
                           None, lp.size, lp.step.data_ptr(), self._actor_learning_rate, 0.9, 0.999, 1e-8, wd, 0.0, 1, st)
            # the reported mean_std is taken after the optimizer step (awac_impl.py:97-99)
            L.gauss_mean_std(lp.params.data_ptr(), self.MIN_LOGSTD, self.MAX_LOGSTD, A, self.metric_ptr(M_STD), st)

    def _sync(self, hard: bool):
        L, st = self._lib, self._stream
        arenas = [self._q_func.arena, self._policy.arena] + ([self._logstd] if self._logstd is not None else [])
        for a in arenas:
            if hard:
                L.hard_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, st)
            else:
                L.soft_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, self._tau, st)
        self._q_func.refresh_shadow("target", st)
        self._policy.refresh_shadow("target", st)

    # ------------------------------------------------------------------ reference hooks (eager, one sync each)
    def _begin(self, batch, *ticks):
        db = self.load_batch(batch)
        if ticks:
            self._tick(*ticks)
        self.zero_slots()
        self.fill_noise(db.B)
        return db

    def update_critic(self, batch) -> np.ndarray:
        db = self._begin(batch, C_DRAW, C_CRITIC)
        self._p_critic(db)
        return self.read_slots()[M_CRITIC].copy()

    def compute_critic_loss(self, batch, q_tpn: torch.Tensor) -> torch.Tensor:
        db = self._begin(batch)
        self._p_critic(db, backward=False, q_tpn=q_tpn.to(self._device).reshape(-1).contiguous())
        self.sync()
        return self._slots[M_CRITIC].clone()

    def compute_actor_loss(self, batch) -> torch.Tensor:
        db = self._begin(batch)
        self._p_actor(db, step=False)
        self.sync()
        if self._logstd is not None:
            self._logstd.grads.zero_()   # the hook only reports the loss
        return self._slots[M_ACTOR].clone()

    def update_critic_target(self) -> None:
        a = self._q_func.arena
        self._lib.soft_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, self._tau, self._stream)
        self._q_func.refresh_shadow("target", self._stream)

    def update_actor_target(self) -> None:
        for a in [self._policy.arena] + ([self._logstd] if self._logstd is not None else []):
            self._lib.soft_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, self._tau, self._stream)
        self._policy.refresh_shadow("target", self._stream)

    # ---- evaluation API
    def _predict_best_action(self, obs: torch.Tensor) -> torch.Tensor:
        """GaussianDistribution.mean = tanh(mu) (policies.py:176-181, distributions.py:83-84)."""
        head = self._policy_head(obs, head_tanh=False)
        return torch.tanh(head[0, :, :self._action_size])

    def sample_action(self, x) -> np.ndarray:
        """dist.sample(): Normal(tanh(mu), std).rsample().clamp(-1, 1) (distributions.py:52-53); host draw."""
        obs = self._eval_obs(x)
        head = self._policy_head(obs, head_tanh=False)
        self.sync()
        h = head[0].detach().cpu().numpy()
        A = self._action_size
        mean = np.tanh(h[:, :A])
        if self.LOGSTD_PARAM:
            p = self._logstd.params[:A].detach().cpu().numpy()
            ls = self.MIN_LOGSTD + (self.MAX_LOGSTD - self.MIN_LOGSTD) / (1.0 + np.exp(-p))
        else:
            ls = np.clip(h[:, A:2 * A], self.MIN_LOGSTD, self.MAX_LOGSTD)
        act = np.clip(mean + np.exp(ls) * np.random.randn(*mean.shape), -1.0, 1.0).astype(np.float32)
        if self._action_scaler is None:
            return act
        with torch.cuda.stream(self._stream_obj):   # action_scaler.reverse_transform (algos/torch/base.py:77-79)
            d = torch.from_numpy(act).to(self._device).contiguous()
        self.unscale_actions(d)
        self.sync()
        return d.cpu().numpy()


class AWACImpl(_GaussActorImpl):
    LOGSTD_PARAM = True
    MIN_LOGSTD, MAX_LOGSTD = -6.0, 0.0     # awac_impl.py:69-77
    MEMBER_REDUCE = 0                      # q_func(x, a, "min")
    ACTOR_WEIGHT_DECAY = 1e-4              # algos/awac.py:105

    def __init__(self, *, lam: float = 1.0, actor_weight_decay: float = 1e-4, **kw):
        super().__init__(**kw)
        self._lam = float(lam)
        self.ACTOR_WEIGHT_DECAY = float(actor_weight_decay)

    def build(self) -> None:
        super().build()
        # AWACImpl is a SACImpl with a frozen temperature (awac_impl.py:46-48: initial_temperature 1e-20,
        # temp_learning_rate 0): the scalar and its never-stepped Adam exist only so that checkpoints interchange
        import math

        from .cql_impl import _Scalar
        from .ddpg_impl import C_TEMP

        self._log_temp = _Scalar(math.log(1e-20), self._device, self._counters[C_TEMP:C_TEMP + 1])

    def _checkpoint_views(self):
        v = super()._checkpoint_views()
        v.update({"_log_temp": self._log_temp, "_temp_optim": self._log_temp.optim_view(0.0)})
        return v

    def _weights_args(self):
        return 0, 0, self._lam, 0.0        # value = mean over samples, softmax over the batch of adv / lam, times B

    def _loss_scale(self, B: int) -> float:
        return 1.0                          # -(log_probs * weights).sum()

    def update_actor(self, batch):
        db = self._begin(batch, C_DRAW, C_ACTOR)
        self._p_actor(db)
        v = self.read_slots()
        return v[M_ACTOR].copy(), v[M_STD].copy()

    def update_fused(self, batch, actor_step: bool):
        return self._metrics_dict(self.update_fused_async(batch, actor_step))

    def update_fused_async(self, batch, actor_step: bool):
        """AWAC._update (algos/awac.py:176-191)."""
        db = self.load_batch(batch, defer=True)

        def program():
            self._tick(C_DRAW, C_CRITIC, *([C_ACTOR] if actor_step else []))
            self.zero_slots()
            self.fill_noise(db.B)
            self._p_critic(db)
            if actor_step:
                self._p_actor(db)
                self._sync(hard=False)

        self.run_program(("awac", db.B, actor_step, self._noise_injected), program)
        return [(M_CRITIC, "critic_loss")] + ([(M_ACTOR, "actor_loss"), (M_STD, "mean_std")] if actor_step else [])


class CRRImpl(_GaussActorImpl):
    LOGSTD_PARAM = False
    MIN_LOGSTD, MAX_LOGSTD = -20.0, 2.0    # create_non_squashed_normal_policy defaults (models/builders.py)
    TARGET_FROM_TARGET_POLICY = True       # crr_impl.py:143-153
    MEMBER_REDUCE = 1                      # q_func(x, a) -> default reduction "mean"

    def __init__(self, *, beta: float = 1.0, advantage_type: str = "mean", weight_type: str = "exp",
                 max_weight: float = 20.0, **kw):
        super().__init__(**kw)
        if advantage_type not in ("mean", "max"):
            raise ValueError(f"invalid advantage type: {advantage_type}.")     # crr_impl.py:139-141
        if weight_type not in ("binary", "exp"):
            raise ValueError(f"invalid weight type: {weight_type}.")           # crr_impl.py:100-102
        self._beta, self._advantage_type, self._weight_type = float(beta), advantage_type, weight_type
        self._max_weight = float(max_weight)

    def _weights_args(self):
        return (0 if self._advantage_type == "mean" else 1, 2 if self._weight_type == "binary" else 1, self._beta,
                self._max_weight)

    def _loss_scale(self, B: int) -> float:
        return 1.0 / B                      # -(log_probs * weight).mean()

    def update_actor(self, batch) -> np.ndarray:
        db = self._begin(batch, C_DRAW, C_ACTOR)
        self._p_actor(db)
        return self.read_slots()[M_ACTOR].copy()

    def sync_critic_target(self) -> None:
        a = self._q_func.arena
        self._lib.hard_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, self._stream)
        self._q_func.refresh_shadow("target", self._stream)

    def sync_actor_target(self) -> None:
        a = self._policy.arena
        self._lib.hard_sync(a.target.data_ptr(), a.params.data_ptr(), a.size, self._stream)
        self._policy.refresh_shadow("target", self._stream)

    def update_fused(self, batch, target_update: str):
        return self._metrics_dict(self.update_fused_async(batch, target_update))

    def update_fused_async(self, batch, target_update: str):
        """CRR._update (algos/crr.py:226-244); target_update in {"hard", "soft", "none"} for this step."""
        db = self.load_batch(batch, defer=True)

        def program():
            self._tick(C_DRAW, C_CRITIC, C_ACTOR)
            self.zero_slots()
            self.fill_noise(db.B)
            self._p_critic(db)
            self._p_actor(db)
            if target_update != "none":
                self._sync(hard=target_update == "hard")

        self.run_program(("crr", db.B, target_update, self._noise_injected), program)
        return [(M_CRITIC, "critic_loss"), (M_ACTOR, "actor_loss")]
