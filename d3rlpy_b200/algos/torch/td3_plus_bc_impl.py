"""TD3+BC on B200: mirrors TD3PlusBCImpl/TD3Impl/DDPGImpl
(d3rlpy/algos/torch/td3_plus_bc_impl.py:16-70, td3_impl.py:15-78, ddpg_impl.py:138-209)."""
from __future__ import annotations

import numpy as np
import torch

from ...nets import DenseNet
from .ddpg_impl import C_ACTOR, C_CRITIC, C_DRAW, DDPGBaseImpl

M_CRITIC, M_ACTOR = 0, 1
S_TD, S_ACT = 0, 4  # sums slots: TD uses [0..2], actor stats [4..6]


class TD3PlusBCImpl(DDPGBaseImpl):
    POLICY_KIND = "deterministic"
    SUPPORTS_QR = True         # ContinuousQRQFunction critics: csrc/qr.cu kernels with a single pseudo-action
    BEHAVIOUR_CLONING = True   # TD3Impl below: plain TD3 actor loss -Q_0(s, pi(s)).mean()

    def __init__(self, *, target_smoothing_sigma=0.2, target_smoothing_clip=0.5, alpha=2.5, **kw):
        super().__init__(**kw)
        self._target_smoothing_sigma = target_smoothing_sigma
        self._target_smoothing_clip = target_smoothing_clip
        self._alpha = alpha

    def _actor_seed(self, q0, a, db, dq, B, A, inv_b):
        """Actor loss value + its gradient seed dQ_0.  TD3+BC (td3_plus_bc_impl.py:64-70): lambda = alpha / mean|Q_0|
        (one extra exchange of the |Q| sums under data parallelism); TD3 (ddpg_impl.py:268-273): -mean Q_0."""
        L, st = self._lib, self._stream
        if self.BEHAVIOUR_CLONING:
            L.td3bc_actor_stats(q0.data_ptr(), a.data_ptr(), A, db.ptr("act"), A, self.sums_ptr(S_ACT), B, A, st)
            self._allreduce(self._slots[32 + S_ACT:32 + S_ACT + 3])
            L.td3bc_actor_seed(self.sums_ptr(S_ACT), self._alpha, inv_b, A, dq.data_ptr(), B, B, 1,
                               self.metric_ptr(M_ACTOR), st)
            return inv_b
        L.neg_mean_seed(q0.data_ptr(), dq.data_ptr(), self.sums_ptr(S_ACT), B, inv_b, st)
        self._allreduce(self._slots[32 + S_ACT:32 + S_ACT + 1])
        L.copy_d2d(self.metric_ptr(M_ACTOR), self.sums_ptr(S_ACT), 4, st)
        return 0.0  # weight of the (a - pi(s))^2 term in the action gradient

    def _build_actor(self) -> None:
        O, A = self._observation_shape[0], self._action_size
        self._policy = DenseNet(O, self._actor_hidden, [("_fc", A)], 1, self._device, trunk_prefix="_encoder.",
                                with_target=True, seed_gen=self._gen, precision=self._precision)

    def noise_layout(self, B):
        return {"target": ("normal", (B, self._action_size))}

    # ------------------------------------------------------------------ program pieces
    def _p_target(self, db):
        """TD3Impl.compute_target (td3_impl.py:61-78) -> q_t[E,B] (min is taken inside critic_loss)."""
        B, O, A, L, st = db.B, db.O, self._action_size, self._lib, self._stream
        a_next = self.ws("tp_a", 1, B, A)
        self._policy.forward("target", db.ptr("next_obs"), O, B, self._policy.ctx("tp", B, 1, False), a_next, st,
                             head_tanh=True)
        xt = self.ws("xt", B, O + A)
        L.concat_rows(db.ptr("next_obs"), O, a_next.data_ptr(), A, self.noise_view("target", B).data_ptr(),
                      self._target_smoothing_sigma, self._target_smoothing_clip, 0.0, xt.data_ptr(), O + A, B, 1, O, A,
                      st)
        _, q_t = self._critic_rows_forward("target", xt, B, "tq", train=False)
        return q_t

    def _p_critic(self, db, q_t=None, q_tpn=None, sync_target=False):
        B, O, A, L, st, E = db.B, db.O, self._action_size, self._lib, self._stream, self._n_critics
        xc = self.ws("xc", B, O + A)
        L.concat_rows(db.ptr("obs"), O, db.ptr("act"), A, None, 0.0, 0.0, 0.0, xc.data_ptr(), O + A, B, 1, O, A, st)
        acts, q = self._critic_rows_forward("params", xc, B, "cq")
        inv_b = 1.0 / (B * self.world_size)
        nq = self._n_quantiles
        if nq:
            # quantile Huber against the quantile vector of the target member with the smallest mean
            # (ContinuousQRQFunction.compute_error / compute_target, ensemble_q_function.py:47-52,177-184)
            if q_tpn is None:
                q_tpn = self.ws("qr_tpn", B, nq)
                L.qr_target(q_t.data_ptr(), B * nq, q_t.data_ptr(), B * nq, q_tpn.data_ptr(), B, 1, nq, E, st)
            dq = self.ws("dq", E, B, nq)
            L.qr_loss(q.data_ptr(), B * nq, q_tpn.data_ptr(), self.ws("qr_action0", B).data_ptr(), db.ptr("rew"),
                      db.ptr("term"), db.ptr("nsteps"), self._gamma, 0.0, dq.data_ptr(), B * nq,
                      self.ws("qr_partials", 2 * B).data_ptr(), self.sums_ptr(S_TD), B, 1, nq, E, inv_b, 0, st)
            self._allreduce(self._slots[32 + S_TD:32 + S_TD + 3])
            L.dcql_finalize(self.sums_ptr(S_TD), inv_b, 0.0, 0, self.metric_ptr(M_CRITIC), st)
            return xc, acts, dq
        dq = self.ws("dq", E, B)
        L.critic_loss(q.data_ptr(), B, q_t.data_ptr() if q_t is not None else None, B, E,
                      q_tpn.data_ptr() if q_tpn is not None else None, db.ptr("rew"), db.ptr("term"),
                      db.ptr("nsteps"), self._gamma, None, None, 0, A, None, 0.0, dq.data_ptr(), B,
                      self.sums_ptr(S_TD), None, B, E, inv_b, 1, st)
        self._allreduce(self._slots[32 + S_TD:32 + S_TD + 3])
        L.cql_finalize(self.sums_ptr(S_TD), None, inv_b, E, 0.0, 0.0, 0, 0, self.metric_ptr(M_CRITIC), None, st)
        return xc, acts, dq

    def _p_critic_step(self, db, xc, acts, dq, sync_target):
        B = db.B
        self._q_func.backward(xc, self._q_func.in_dim, B, acts, dq, self._stream)
        self._allreduce(self._q_func.arena.grads)
        self._q_func.adam(self._critic_learning_rate, self._stream, tau=self._tau if sync_target else None)

    def _p_actor(self, db, step=True):
        """compute_actor_loss (td3_plus_bc_impl.py:64-70) + backward + Adam; only member 0 is evaluated; step=False
        stops after the loss value."""
        B, O, A, L, st = db.B, db.O, self._action_size, self._lib, self._stream
        acts_p = self._policy.ctx("pi", B, 1, True)
        a = self.ws("pi_a", 1, B, A)
        self._policy.forward("params", db.ptr("obs"), O, B, acts_p, a, st, head_tanh=True)
        xa = self.ws("xa", B, O + A)
        L.concat_rows(db.ptr("obs"), O, a.data_ptr(), A, None, 0.0, 0.0, 0.0, xa.data_ptr(), O + A, B, 1, O, A, st)
        acts_c, q0 = self._critic_rows_forward("params", xa, B, "aq", members=1)
        inv_b = 1.0 / (B * self.world_size)
        nq = self._n_quantiles
        if nq:   # Q_0 = mean of the quantiles (qr_q_function.py:118-122)
            theta0, q0 = q0, self.ws("aq_values", 1, B)
            L.qr_values(theta0.data_ptr(), B * nq, q0.data_ptr(), B, B, 1, nq, 1, st)
        dq = self.ws("a_dq", 1, B)
        bc_w = self._actor_seed(q0, a, db, dq, B, A, inv_b)
        if not step:
            return
        if nq:
            dq, dvalues = self.ws("a_dtheta", 1, B, nq), dq
            L.qr_values_backward(dvalues.data_ptr(), dq.data_ptr(), B, nq, st)
        dxa = self.ws("a_dx", B, A)
        self._q_func.backward(xa, O + A, B, acts_c, dq, st, weight_grads=False, dx=dxa, lddx=A, stride_dx=B * A,
                              dx_col0=O, dx_cols=A)
        dz = self.ws("pi_dz", 1, B, A)
        L.td3bc_actor_backward(a.data_ptr(), A, db.ptr("act"), A, dxa.data_ptr(), A, dz.data_ptr(), A, B, A, bc_w, st)
        self._policy.backward(db.ptr("obs"), O, B, acts_p, dz, st)
        self._allreduce(self._policy.arena.grads)
        self._policy.adam(self._actor_learning_rate, st, tau=self._tau)

    def _allreduce(self, t):
        if self.world_size > 1:
            from ...parallel import allreduce_sum

            allreduce_sum(t, self._stream_obj)

    # ------------------------------------------------------------------ single-GPU tensor-core program
    fused_glue = True

    def _side_stream(self) -> int:
        if getattr(self, "_side_obj", None) is None:
            self._side_obj = torch.cuda.Stream(device=self._device)
        return self._side_obj.cuda_stream

    def _side_stream2(self) -> int:
        if getattr(self, "_side2_obj", None) is None:
            self._side2_obj = torch.cuda.Stream(device=self._device)
        return self._side2_obj.cuda_stream

    def _program_fused(self, db, actor_step: bool):
        """Same update as `program` with the rows written straight as bf16 GEMM operands, the loss tail fused, and
        the online-critic forward running on a graph branch beside the target path (policy' -> smoothing -> Q')."""
        B, O, A, E = db.B, db.O, self._action_size, self._n_critics
        L, st = self._lib, self._stream
        f32 = self._precision == "fp32"   # fp32 mode: fp32 operand rows, one GEMM launch per layer (3xTF32 / SIMT)
        ld = (O + A + 3) // 4 * 4 if f32 else (O + A + 7) // 8 * 8
        esz = 4 if f32 else 2
        inv_b = 1.0 / B
        mask = (1 << C_DRAW) | (1 << C_CRITIC) | ((1 << C_ACTOR) if actor_step else 0)
        # ONE prologue launch: counters, slot zeroing, the update's noise, bf16 operand rows of [obs; next_obs] (the
        # inputs of pi(s) and pi'(s'))
        _, n_norm, n_uni, _ = self._noise_plan(B)
        draw = not self._noise_injected and n_norm + n_uni > 0
        assert db.off["next_obs"] == db.off["obs"] + B * O, "obs/next_obs must be contiguous"
        ldo = (O + 7) // 8 * 8
        xb = None if f32 else self.ws("pi_xb", 2 * B, ldo, dtype=torch.bfloat16)
        L.update_prologue(self._counters.data_ptr(), self.N_COUNTERS, mask, C_DRAW, self._slots.data_ptr(), 64,
                          self._noise_arena(B).data_ptr() if draw else None, n_norm if draw else 0, n_uni if draw else 0,
                          (self._seed + 0x9E3779B97F4A7C15 * self.rank) & 0xFFFFFFFFFFFFFFFF,
                          db.ptr("obs") if xb is not None else None, O, 2 * B, O,
                          xb.data_ptr() if xb is not None else None, ldo, self.ws("pro_done", 4, dtype=torch.int32).data_ptr(),
                          st)
        xb_obs = None if xb is None else (xb.data_ptr(), ldo)
        xb_next = None if xb is None else (xb.data_ptr() + 2 * B * ldo, ldo)
        X = self.ws("xf_rows", 3 * B, ld, dtype=torch.float32 if f32 else torch.bfloat16)   # [critic | target | actor] rows
        concat = L.concat_rows if f32 else L.concat_rows_bf16

        def q_forward(which, xptr, ctx, out, stream):
            if f32:
                q_net.forward(which, xptr, ld, B, ctx, out, stream)
            else:
                q_net.forward(which, None, 0, B, ctx, out, stream, x_bf16=(xptr, ld))

        def q_backward(xptr, ctx, dq_, stream, **kw):
            q_net.backward(xptr if f32 else None, ld if f32 else 0, B, ctx, dq_, stream, **kw)

        done = self.ws("xf_done", 4 + 3 * ((B * E + 7) // 8) + 4, dtype=torch.int32)
        q_net, pi = self._q_func, self._policy
        # ---- branch: online critics on (s, a)
        side = self._side_stream()
        L.stream_fork(st, side)
        concat(db.ptr("obs"), O, db.ptr("act"), A, None, 0.0, 0.0, 0.0, X.data_ptr(), ld, B, 1, O, A, side)
        ctx_c = q_net.ctx("cq", B, E, True)
        q = self.ws("cq_q", E, B)
        q_forward("params", X.data_ptr(), ctx_c, q, side)
        if actor_step:
            # ---- second branch: pi(s) and the actor rows do not depend on the critic step -> beside it, not after it
            side2 = self._side_stream2()
            L.stream_fork(st, side2)
            acts_p = pi.ctx("pi", B, 1, True)
            a = self.ws("pi_a", 1, B, A)
            pi.forward("params", db.ptr("obs"), O, B, acts_p, a, side2, head_tanh=True, x_bf16=xb_obs)
            xa = X.data_ptr() + esz * 2 * B * ld
            concat(db.ptr("obs"), O, a.data_ptr(), A, None, 0.0, 0.0, 0.0, xa, ld, B, 1, O, A, side2)
        # ---- main: target policy -> smoothed action -> target critics
        a_next = self.ws("tp_a", 1, B, A)
        pi.forward("target", db.ptr("next_obs"), O, B, pi.ctx("tp", B, 1, False), a_next, st, head_tanh=True,
                   x_bf16=xb_next)
        xt = X.data_ptr() + esz * B * ld
        concat(db.ptr("next_obs"), O, a_next.data_ptr(), A, self.noise_view("target", B).data_ptr(),
               self._target_smoothing_sigma, self._target_smoothing_clip, 0.0, xt, ld, B, 1, O, A, st)
        q_t = self.ws("tq_q", E, B)
        q_forward("target", xt, q_net.ctx("tq", B, E, False), q_t, st)
        L.stream_join(st, side)
        dq = self.ws("dq", E, B)
        L.cql_loss_step(q.data_ptr(), B, q_t.data_ptr(), B, E, None, db.ptr("rew"), db.ptr("term"), db.ptr("nsteps"),
                        self._gamma, None, None, 0, A, None, 0.0, 0.0, dq.data_ptr(), B, self.sums_ptr(S_TD),
                        done.data_ptr(), B, E, inv_b, 0, None, 0.0, self.metric_ptr(M_CRITIC), None, st)
        q_backward(X.data_ptr(), ctx_c, dq, st)
        q_net.adam(self._critic_learning_rate, st, tau=self._tau if actor_step else None)
        if not actor_step:
            return
        # ---- actor step (td3_plus_bc_impl.py:64-70): member 0 only, on the updated critics
        L.stream_join(st, side2)
        ctx_a = q_net.ctx("aq", B, 1, True)
        q0 = self.ws("aq_q", 1, B)
        q_forward("params", xa, ctx_a, q0, st)
        dq0 = self.ws("a_dq", 1, B)
        bc_w = self._actor_seed(q0, a, db, dq0, B, A, inv_b)
        dxa = self.ws("a_dx", B, A)
        q_backward(xa, ctx_a, dq0, st, weight_grads=False, dx=dxa, lddx=A, stride_dx=B * A, dx_col0=O, dx_cols=A)
        dz = self.ws("pi_dz", 1, B, A)
        L.td3bc_actor_backward(a.data_ptr(), A, db.ptr("act"), A, dxa.data_ptr(), A, dz.data_ptr(), A, B, A, bc_w, st)
        pi.backward(db.ptr("obs"), O, B, acts_p, dz, st)
        pi.adam(self._actor_learning_rate, st, tau=self._tau)

    # ------------------------------------------------------------------ fused update (TD3PlusBC._update)
    def update_fused(self, batch, actor_step: bool):
        return self._metrics_dict(self.update_fused_async(batch, actor_step))

    def update_fused_async(self, batch, actor_step: bool):
        """Enqueue one whole update (no host sync); returns the metric slot names."""
        db = self.load_batch(batch, defer=True)

        def program():
            self._tick(C_DRAW, C_CRITIC, *([C_ACTOR] if actor_step else []))
            self.zero_slots()
            self.fill_noise(db.B)
            q_t = self._p_target(db)
            xc, acts, dq = self._p_critic(db, q_t=q_t)
            self._p_critic_step(db, xc, acts, dq, sync_target=actor_step)
            if actor_step:
                self._p_actor(db)

        fused = (self.world_size == 1 and self.fused_glue and not self._n_quantiles
                 and ((self._precision == "bf16" and self._q_func.fused_ok and self._policy.fused_ok)
                      or (self._precision == "fp32" and not self._q_func.wide_head)))
        self.run_program(("td3bc", db.B, actor_step, self._noise_injected, fused),
                         (lambda: self._program_fused(db, actor_step)) if fused else program)
        return [(M_CRITIC, "critic_loss")] + ([(M_ACTOR, "actor_loss")] if actor_step else [])

    # ------------------------------------------------------------------ reference hooks (eager)
    def compute_target(self, batch) -> torch.Tensor:
        db = self.load_batch(batch)
        self.fill_noise(db.B)
        q_t = self._p_target(db)
        nq = self._n_quantiles
        if nq:   # (B, n_quantiles) of the member with the smallest mean (ensemble_q_function.py:47-52)
            q_tpn = self.ws("qr_tpn", db.B, nq)
            self._lib.qr_target(q_t.data_ptr(), db.B * nq, q_t.data_ptr(), db.B * nq, q_tpn.data_ptr(), db.B, 1, nq,
                                self._n_critics, self._stream)
            self.sync()
            return q_tpn.clone()
        self.sync()
        return q_t.min(dim=0).values.view(-1, 1).clone()

    def compute_critic_loss(self, batch, q_tpn: torch.Tensor) -> torch.Tensor:
        db = self.load_batch(batch)
        self.zero_slots()
        self._p_critic(db, q_tpn=q_tpn.to(self._device).reshape(-1).contiguous())
        self.sync()
        return self._slots[M_CRITIC].clone()

    def update_critic(self, batch) -> np.ndarray:
        db = self.load_batch(batch)
        self._tick(C_DRAW, C_CRITIC)
        self.zero_slots()
        self.fill_noise(db.B)
        q_t = self._p_target(db)
        xc, acts, dq = self._p_critic(db, q_t=q_t)
        self._p_critic_step(db, xc, acts, dq, sync_target=False)
        return self.read_slots()[M_CRITIC].copy()

    def compute_actor_loss(self, batch) -> torch.Tensor:
        """TD3PlusBCImpl.compute_actor_loss (td3_plus_bc_impl.py:64-70): -lambda * mean Q_0(s, pi(s)) + mean((a - pi(s))^2)
        with lambda = alpha / mean|Q_0| detached; TD3 / DDPG: -mean Q_0(s, pi(s)) (ddpg_impl.py:268-273).  Nothing is
        stepped."""
        db = self.load_batch(batch)
        self.zero_slots()
        self._p_actor(db, step=False)
        self.sync()
        return self._slots[M_ACTOR].clone()

    def update_actor(self, batch) -> np.ndarray:
        db = self.load_batch(batch)
        self._tick(C_ACTOR)
        self.zero_slots()
        self._p_actor_no_sync(db)
        return self.read_slots()[M_ACTOR].copy()

    def _p_actor_no_sync(self, db):
        tau, self._tau = self._tau, None
        try:
            self._p_actor(db)
        finally:
            self._tau = tau

    # ---- evaluation API
    def _predict_best_action(self, obs: torch.Tensor) -> torch.Tensor:
        """DeterministicPolicy.best_action = tanh(fc(encoder(x))) (policies.py:57-79)."""
        return self._policy_head(obs, head_tanh=True)[0]

    def sample_action(self, x) -> np.ndarray:
        return self.predict_best_action(x)


class TD3Impl(TD3PlusBCImpl):
    """TD3Impl (d3rlpy/algos/torch/td3_impl.py:15-78): the same target smoothing, critic step and delayed actor step;
    the actor loss is DDPGImpl's -Q_0(s, pi(s)).mean() (ddpg_impl.py:268-273)."""

    BEHAVIOUR_CLONING = False

    def __init__(self, **kw):
        kw.pop("alpha", None)
        super().__init__(alpha=0.0, **kw)


class DDPGImpl(TD3Impl):
    """DDPGImpl (d3rlpy/algos/torch/ddpg_impl.py:255-288): TD3Impl with `target_smoothing_sigma = 0` — the smoothing
    term `clamp(0 * noise, -c, c)` vanishes and the target action is `clamp(pi'(s'), -1, 1)` (ddpg_impl.py:279-284)."""
