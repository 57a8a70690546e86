"""BCQ on B200: mirrors BCQImpl (d3rlpy/algos/torch/bcq_impl.py:27-226) over DDPGBaseImpl.

Networks: critic ensemble (mean Q), perturbation policy `DeterministicResidualPolicy`
(policies.py:82-105) and the `ConditionalVAE` imitator (imitators.py:13-118).  The VAE's two
trunks live in two arenas that share ONE Adam step counter (Adam is element-wise, so this equals the
reference's single `imitator_optim`); `impl.imitator.state_dict()` exposes the reference's key order.

One update (BCQ._update, d3rlpy/algos/bcq.py:261-279):
  imitator: enc fwd -> sample z -> dec fwd -> MSE + beta KL -> dec bwd (dW + d latent) -> enc bwd -> Adam
  critic  : B*N rows [s' | clamp(randn)] -> decoder -> target perturbation policy -> target critics
            -> max_k((1-lam) max_e + lam min_e) -> TD loss -> bwd -> Adam
  actor   : [s | clamp(randn)] -> decoder (no grad kept) -> policy -> Q_0 -> -mean ; dgrad through Q_0 to
            the action, clamp/tanh gates, policy bwd -> Adam ; soft syncs (actor target, then critic target).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
from typing import Optional
import torch

from ...nets import DenseNet
from .ddpg_impl import C_ACTOR, C_CRITIC, C_DRAW, C_IMITATOR, DDPGBaseImpl, _ModuleView

M_IMITATOR, M_CRITIC, M_ACTOR = 0, 1, 2
S_VAE, S_TD, S_ACT = 0, 4, 8
VAE_MIN_LOGSTD, VAE_MAX_LOGSTD = -4.0, 15.0  # create_conditional_vae defaults (models/builders.py:197-222)


class _ImitatorView:
    """`impl.imitator`-style view: reference key order encoder trunk, decoder trunk, _mu, _logstd, _fc."""

    def __init__(self, enc: DenseNet, dec: DenseNet):
        self._enc, self._dec = enc, dec

    def state_dict(self, which: str = "params"):
        e, d = self._enc.arena.state_dict(which), self._dec.arena.state_dict(which)
        out = OrderedDict()
        for k, v in e.items():
            if k.startswith("_encoder_encoder."):
                out[k] = v
        for k, v in d.items():
            if k.startswith("_decoder_encoder."):
                out[k] = v
        for k in ("_mu.weight", "_mu.bias", "_logstd.weight", "_logstd.bias"):
            out[k] = e[k]
        for k in ("_fc.weight", "_fc.bias"):
            out[k] = d[k]
        return out

    def load_state_dict(self, sd):
        for net in (self._enc, self._dec):
            mine = net.arena.state_dict()
            net.arena.load_state_dict({k: sd[k] for k in mine})
            st = torch.cuda.current_stream(net.device)
            net.refresh_shadow("params", st.cuda_stream)
            st.synchronize()

    def parameters(self):
        return list(self.state_dict().values())


class BCQImpl(DDPGBaseImpl):
    POLICY_KIND = "bcq"
    def __init__(self, *, imitator_learning_rate=1e-3, imitator_hidden=(750, 750), lam=0.75, n_action_samples=100,
                 action_flexibility=0.05, beta=0.5, **kw):
        super().__init__(**kw)
        self._imitator_learning_rate = imitator_learning_rate
        self._imitator_hidden = list(imitator_hidden)
        self._lam, self._n_action_samples = lam, n_action_samples
        self._action_flexibility, self._beta = action_flexibility, beta

    def _build_actor(self) -> None:
        O, A = self._observation_shape[0], self._action_size
        self._policy = DenseNet(O + A, self._actor_hidden, [("_fc", A)], 1, self._device, trunk_prefix="_encoder.",
                                with_target=True, seed_gen=self._gen, precision=self._precision)

    def build(self) -> None:
        super().build()
        O, A = self._observation_shape[0], self._action_size
        Lz = 2 * A
        self._vae_enc = DenseNet(O + A, self._imitator_hidden, [("_mu", Lz), ("_logstd", Lz)], 1, self._device,
                                 trunk_prefix="_encoder_encoder.", seed_gen=self._gen, precision=self._precision)
        self._vae_dec = DenseNet(O + Lz, self._imitator_hidden, [("_fc", A)], 1, self._device,
                                 trunk_prefix="_decoder_encoder.", seed_gen=self._gen, precision=self._precision)
        for net in (self._vae_enc, self._vae_dec):
            net.arena.step = self._counters[C_IMITATOR:C_IMITATOR + 1]
            net.refresh_shadow("params", self._stream)
        self.sync()

    @property
    def imitator(self):
        return _ImitatorView(self._vae_enc, self._vae_dec)

    def _checkpoint_views(self):
        from .ddpg_impl import _OptimView

        v = super()._checkpoint_views()
        im = self.imitator
        v.update({"_imitator": im,
                  "_imitator_optim": _OptimView(lambda which: im.state_dict(which), self._vae_enc.arena.step,
                                                self._imitator_learning_rate)})
        return v

    def noise_layout(self, B):
        """Draw order: imitator eps (B,2A); critic randn(B*N,2A); actor randn(B,2A)."""
        A, N = self._action_size, self._n_action_samples
        return {"imitator": ("normal", (B, 2 * A)), "critic": ("normal", (B * N, 2 * A)),
                "actor": ("normal", (B, 2 * A))}

    # ------------------------------------------------------------------ program pieces
    def _p_imitator(self, db):
        """update_imitator (bcq_impl.py:148-161) = ConditionalVAE.compute_error (imitators.py:80-86)."""
        B, O, A, L, st = db.B, db.O, self._action_size, self._lib, self._stream
        Lz = 2 * A
        enc, dec = self._vae_enc, self._vae_dec
        inv_b = 1.0 / (B * self.world_size)
        xe = self.ws("vae_xe", B, O + A)
        L.concat_rows(db.ptr("obs"), O, db.ptr("act"), A, None, 0.0, 0.0, 0.0, xe.data_ptr(), O + A, B, 1, O, A, st)
        ce = enc.ctx("vae_e", B, 1, True)
        head = self.ws("vae_head", 1, B, 2 * Lz)
        enc.forward("params", xe, O + A, B, ce, head, st)
        xd = self.ws("vae_xd", B, O + Lz)
        eps = self.noise_view("imitator", B)
        L.vae_sample_rows(head.data_ptr(), 2 * Lz, eps.data_ptr(), db.ptr("obs"), O, xd.data_ptr(), O + Lz,
                          self.sums_ptr(S_VAE), B, O, Lz, VAE_MIN_LOGSTD, VAE_MAX_LOGSTD, st)
        cd = dec.ctx("vae_d", B, 1, True)
        y = self.ws("vae_y", 1, B, A)
        dec.forward("params", xd, O + Lz, B, cd, y, st, head_tanh=True)
        dpre = self.ws("vae_dpre", 1, B, A)
        L.vae_recon(y.data_ptr(), db.ptr("act"), A, dpre.data_ptr(), self.sums_ptr(S_VAE + 1), B, A, inv_b, st)
        dz = self.ws("vae_dz", 1, B, Lz)
        dec.backward(xd, O + Lz, B, cd, dpre, st, dx=dz, lddx=Lz, stride_dx=B * Lz, dx_col0=O, dx_cols=Lz)
        dhead = self.ws("vae_dhead", 1, B, 2 * Lz)
        L.vae_backward(head.data_ptr(), 2 * Lz, eps.data_ptr(), dz.data_ptr(), Lz, dhead.data_ptr(), 2 * Lz, B, Lz,
                       VAE_MIN_LOGSTD, VAE_MAX_LOGSTD, self._beta, inv_b, st)
        enc.backward(xe, O + A, B, ce, dhead, st)
        self._allreduce(self._slots[32 + S_VAE:32 + S_VAE + 2])
        L.vae_finalize(self.sums_ptr(S_VAE), A, Lz, self._beta, inv_b, self.metric_ptr(M_IMITATOR), st)
        for net in (enc, dec):
            self._allreduce(net.arena.grads)
            net.adam(self._imitator_learning_rate, st)

    def _p_target(self, db):
        """compute_target (bcq_impl.py:163-187,215-226) + compute_max_with_n_actions
        (q_functions/__init__.py:8-63) -> q_tpn[B]."""
        B, O, A, N, L, st = db.B, db.O, self._action_size, self._n_action_samples, self._lib, self._stream
        Lz, R = 2 * A, db.B * self._n_action_samples
        xd = self.ws("t_xd", R, O + Lz)
        L.concat_rows(db.ptr("next_obs"), O, self.noise_view("critic", B).data_ptr(), Lz, None, 0.0, 0.0, 0.5,
                      xd.data_ptr(), O + Lz, B, N, O, Lz, st)
        sampled = self.ws("t_sampled", 1, R, A)
        self._vae_dec.forward("params", xd, O + Lz, R, self._vae_dec.ctx("t_d", R, 1, False), sampled, st,
                              head_tanh=True)
        xp = self.ws("t_xp", R, O + A)
        L.concat_rows(db.ptr("next_obs"), O, sampled.data_ptr(), A, None, 0.0, 0.0, 0.0, xp.data_ptr(), O + A, B, N,
                      O, A, st)
        z = self.ws("t_z", 1, R, A)
        self._policy.forward("target", xp, O + A, R, self._policy.ctx("t_p", R, 1, False), z, st)
        xq = self.ws("t_xq", R, O + A)
        L.residual_rows(z.data_ptr(), A, sampled.data_ptr(), A, db.ptr("next_obs"), O, xq.data_ptr(), O + A,
                        self._action_flexibility, R, N, O, A, st)
        _, q = self._critic_rows_forward("target", xq, R, "t_q", train=False)
        q_tpn = self.ws("q_tpn", B)
        L.bcq_target_reduce(q.data_ptr(), R, q_tpn.data_ptr(), B, N, self._n_critics, self._lam, st)
        return q_tpn

    # ------------------------------------------------------------------ evaluation (bcq_impl.py:163-211)
    def _predict_best_action(self, obs: torch.Tensor, latent: Optional[torch.Tensor] = None) -> torch.Tensor:
        """BCQImpl._predict_best_action: N candidate actions per observation from the decoder on clamp(randn, +-0.5),
        perturbed by the (online) residual policy, scored by critic 0; returns the arg-max candidate.
        `latent` ([n*N, 2A], row = observation-major) injects the draws (parity tests); otherwise Philox."""
        n, O, A, N, L, st = obs.shape[0], obs.shape[1], self._action_size, self._n_action_samples, self._lib, self._stream
        Lz, R = 2 * A, n * N
        eps = self.ws("e_eps", R, Lz)
        if latent is not None:
            eps.copy_(latent.to(self._device).reshape(R, Lz))
        else:
            self._tick(0)
            seed = (self._seed + 0x51ED270B * (self.rank + 1)) & 0xFFFFFFFFFFFFFFFF
            L.noise_fill(eps.data_ptr(), R * Lz, 0, seed, self.counter_ptr(0), st)
        xd = self.ws("e_xd", R, O + Lz)
        L.concat_rows(obs.data_ptr(), O, eps.data_ptr(), Lz, None, 0.0, 0.0, 0.5, xd.data_ptr(), O + Lz, n, N, O, Lz, st)
        sampled = self.ws("e_sampled", 1, R, A)
        self._vae_dec.forward("params", xd, O + Lz, R, self._vae_dec.ctx("e_d", R, 1, False), sampled, st, head_tanh=True)
        xp = self.ws("e_xp", R, O + A)
        L.concat_rows(obs.data_ptr(), O, sampled.data_ptr(), A, None, 0.0, 0.0, 0.0, xp.data_ptr(), O + A, n, N, O, A, st)
        z = self.ws("e_z", 1, R, A)
        self._policy.forward("params", xp, O + A, R, self._policy.ctx("e_p", R, 1, False), z, st)
        xq = self.ws("e_xq", R, O + A)
        L.residual_rows(z.data_ptr(), A, sampled.data_ptr(), A, obs.data_ptr(), O, xq.data_ptr(), O + A,
                        self._action_flexibility, R, N, O, A, st)
        _, q = self._critic_rows_forward("params", xq, R, "e_q", train=False)
        with torch.cuda.stream(self._stream_obj):
            index = q[0, :R].view(n, N).argmax(dim=1)
            return xq[:, O:O + A].reshape(n, N, A)[torch.arange(n, device=self._device), index].clone()

    def sample_action(self, x):
        raise NotImplementedError("BCQ does not support sampling action")  # bcq_impl.py:213-214

    def _p_critic_forward(self, db, stream=None):
        """Online critics on (s, a): independent of the imitator step and of the target computation."""
        B, O, A, L = db.B, db.O, self._action_size, self._lib
        st = self._stream if stream is None else stream
        xc = self.ws("xc", B, O + A)
        L.concat_rows(db.ptr("obs"), O, db.ptr("act"), A, None, 0.0, 0.0, 0.0, xc.data_ptr(), O + A, B, 1, O, A, st)
        acts, q = self._critic_rows_forward("params", xc, B, "cq", stream=st)
        return xc, acts, q

    def _p_critic(self, db, q_tpn, step=True, sync_target=False, fwd=None):
        B, O, A, L, st, E = db.B, db.O, self._action_size, self._lib, self._stream, self._n_critics
        xc, acts, q = fwd if fwd is not None else self._p_critic_forward(db)
        dq = self.ws("dq", E, B)
        inv_b = 1.0 / (B * self.world_size)
        L.critic_loss(q.data_ptr(), B, None, B, E, q_tpn.data_ptr(), db.ptr("rew"), db.ptr("term"), db.ptr("nsteps"),
                      self._gamma, None, None, 0, A, None, 0.0, dq.data_ptr(), B, self.sums_ptr(S_TD), None, B, E,
                      inv_b, 1, st)
        self._allreduce(self._slots[32 + S_TD:32 + S_TD + 3])
        L.cql_finalize(self.sums_ptr(S_TD), None, inv_b, E, 0.0, 0.0, 0, 0, self.metric_ptr(M_CRITIC), None, st)
        if step:
            self._q_func.backward(xc, O + A, B, acts, dq, st)
            self._allreduce(self._q_func.arena.grads)
            self._q_func.adam(self._critic_learning_rate, st, tau=self._tau if sync_target else None)

    def _p_actor(self, db, sync_target=True, step=True, front=None):
        """compute_actor_loss (bcq_impl.py:132-146): -Q_0(s, pi(s, decode(s, clamp(randn)))).mean(); step=False stops
        after the loss value."""
        front = front if front is not None else self._p_actor_front(db)
        self._p_actor_back(db, front, sync_target, step)

    def _p_actor_front(self, db, stream=None):
        """decode(s, clamp(randn)) -> residual policy -> critic input rows: needs the updated imitator and the current
        policy, not the critic step."""
        B, O, A, L = db.B, db.O, self._action_size, self._lib
        st = self._stream if stream is None else stream
        Lz = 2 * A
        xd = self.ws("a_xd", B, O + Lz)
        L.concat_rows(db.ptr("obs"), O, self.noise_view("actor", B).data_ptr(), Lz, None, 0.0, 0.0, 0.5,
                      xd.data_ptr(), O + Lz, B, 1, O, Lz, st)
        sampled = self.ws("a_sampled", 1, B, A)
        self._vae_dec.forward("params", xd, O + Lz, B, self._vae_dec.ctx("a_d", B, 1, False), sampled, st,
                              head_tanh=True)
        xp = self.ws("a_xp", B, O + A)
        L.concat_rows(db.ptr("obs"), O, sampled.data_ptr(), A, None, 0.0, 0.0, 0.0, xp.data_ptr(), O + A, B, 1, O, A,
                      st)
        cp = self._policy.ctx("pi", B, 1, True)
        z = self.ws("a_z", 1, B, A)
        self._policy.forward("params", xp, O + A, B, cp, z, st)
        xq = self.ws("a_xq", B, O + A)
        L.residual_rows(z.data_ptr(), A, sampled.data_ptr(), A, db.ptr("obs"), O, xq.data_ptr(), O + A,
                        self._action_flexibility, B, 1, O, A, st)
        return xp, cp, z, sampled, xq

    def _p_actor_back(self, db, front, sync_target=True, step=True):
        B, O, A, L, st = db.B, db.O, self._action_size, self._lib, self._stream
        xp, cp, z, sampled, xq = front
        cc, q0 = self._critic_rows_forward("params", xq, B, "aq", members=1)
        dq = self.ws("a_dq", 1, B)
        inv_b = 1.0 / (B * self.world_size)
        L.neg_mean_seed(q0.data_ptr(), dq.data_ptr(), self.sums_ptr(S_ACT), B, inv_b, st)
        self._allreduce(self._slots[32 + S_ACT:32 + S_ACT + 1])
        L.copy_d2d(self.metric_ptr(M_ACTOR), self.sums_ptr(S_ACT), 4, st)
        if not step:
            return
        da = self.ws("a_da", B, A)
        self._q_func.backward(xq, O + A, B, cc, dq, st, weight_grads=False, dx=da, lddx=A, stride_dx=B * A,
                              dx_col0=O, dx_cols=A)
        dz = self.ws("a_dz", 1, B, A)
        L.residual_backward(z.data_ptr(), A, sampled.data_ptr(), A, da.data_ptr(), A, dz.data_ptr(), A,
                            self._action_flexibility, B, A, st)
        self._policy.backward(xp, O + A, B, cp, dz, st)
        self._allreduce(self._policy.arena.grads)
        self._policy.adam(self._actor_learning_rate, st, tau=self._tau if sync_target else None)

    def _side_streams(self):
        if getattr(self, "_side_objs", None) is None:
            self._side_objs = (torch.cuda.Stream(device=self._device), torch.cuda.Stream(device=self._device))
        return self._side_objs[0].cuda_stream, self._side_objs[1].cuda_stream

    def _allreduce(self, t):
        if self.world_size > 1:
            from ...parallel import allreduce_sum

            allreduce_sum(t, self._stream_obj)

    # ------------------------------------------------------------------ fused update (BCQ._update)
    def update_fused(self, batch, rl_step: bool, actor_step: bool):
        return self._metrics_dict(self.update_fused_async(batch, rl_step, actor_step))

    def update_fused_async(self, batch, rl_step: bool, actor_step: bool):
        db = self.load_batch(batch, defer=True)
        actor_step = actor_step and rl_step

        def program():
            ticks = [C_DRAW, C_IMITATOR] + ([C_CRITIC] if rl_step else []) + ([C_ACTOR] if actor_step else [])
            self._tick(*ticks)
            self.zero_slots()
            self.fill_noise(db.B)
            L, st = self._lib, self._stream
            fwd = front = None
            if rl_step:
                # graph branch 1: the online critics on (s, a) depend on nothing else in the update
                s1, s2 = self._side_streams()
                L.stream_fork(st, s1)
                fwd = self._p_critic_forward(db, stream=s1)
            self._p_imitator(db)
            if rl_step:
                if actor_step:
                    # graph branch 2: the front of the actor step needs the updated imitator only
                    L.stream_fork(st, s2)
                    front = self._p_actor_front(db, stream=s2)
                q_tpn = self._p_target(db)
                # reference order: critic step, actor step, actor-target sync, critic-target sync
                # (bcq.py:270-277); the critic target depends only on the critic params, so its soft
                # sync is fused into the critic Adam pass.
                L.stream_join(st, s1)
                self._p_critic(db, q_tpn, sync_target=actor_step, fwd=fwd)
                if actor_step:
                    L.stream_join(st, s2)
                    self._p_actor(db, front=front)

        self.run_program(("bcq", db.B, rl_step, actor_step, self._noise_injected), program)
        names = [(M_IMITATOR, "imitator_loss")]
        if rl_step:
            names.append((M_CRITIC, "critic_loss"))
        if actor_step:
            names.append((M_ACTOR, "actor_loss"))
        return names

    # ------------------------------------------------------------------ reference hooks (eager)
    def _begin(self, batch, *ticks):
        db = self.load_batch(batch)
        if ticks:
            self._tick(*ticks)
        self.zero_slots()
        self.fill_noise(db.B)
        return db

    def update_imitator(self, batch) -> np.ndarray:
        db = self._begin(batch, C_DRAW, C_IMITATOR)
        self._p_imitator(db)
        return self.read_slots()[M_IMITATOR].copy()

    def compute_target(self, batch) -> torch.Tensor:
        db = self._begin(batch)
        q = self._p_target(db)
        self.sync()
        return q.view(-1, 1).clone()

    def compute_critic_loss(self, batch, q_tpn: torch.Tensor) -> torch.Tensor:
        db = self._begin(batch)
        self._p_critic(db, q_tpn.to(self._device).reshape(-1).contiguous(), step=False)
        self.sync()
        return self._slots[M_CRITIC].clone()

    def update_critic(self, batch) -> np.ndarray:
        db = self._begin(batch, C_DRAW, C_CRITIC)
        self._p_critic(db, self._p_target(db))
        return self.read_slots()[M_CRITIC].copy()

    def compute_actor_loss(self, batch) -> torch.Tensor:
        """BCQImpl.compute_actor_loss (bcq_impl.py:132-146); nothing is stepped."""
        db = self._begin(batch)
        self._p_actor(db, step=False)
        self.sync()
        return self._slots[M_ACTOR].clone()

    def update_actor(self, batch) -> np.ndarray:
        db = self._begin(batch, C_DRAW, C_ACTOR)
        self._p_actor(db, sync_target=False)
        return self.read_slots()[M_ACTOR].copy()
