"""BEAR on B200: mirrors BEARImpl (d3rlpy/algos/torch/bear_impl.py:41-329), which extends SACImpl with the conditional
VAE of BCQ (imitators.py:13-118, min / max logstd -4 / 15), a Lagrange multiplier log_alpha on the MMD constraint
(clamped to [-5, 10]) and a BCQ-like target.  Built over CQLImpl's SAC pieces (policy pass, temperature step, SAC actor
loss) and BCQImpl's VAE step.  One update (BEAR._update, algos/bear.py:279-309):

  imitator step -> temperature step -> alpha step (MMD on fresh samples, no policy gradient) -> critic step against
  the best of n_target policy actions (lam-mix over the target members) minus the entropy term of that action ->
  actor step on the MMD loss alone during warm-up, on SAC's loss + the MMD loss afterwards -> actor / critic soft syncs.

Loss tails: csrc/bear.cu.  Reference draw order per update: VAE eps (B, 2A); temperature eps (B, A); alpha-step decoder
latents (n B, 2A) and policy eps (n, B, A); target eps (n_target, B, A); [SAC actor eps (B, A) after warm-up]; actor-step
decoder latents and policy eps."""
from __future__ import annotations

import math

import numpy as np
import torch

from ...nets import DenseNet
from .bcq_impl import VAE_MAX_LOGSTD, VAE_MIN_LOGSTD, _ImitatorView
from .cql_impl import MAX_LOGSTD, MIN_LOGSTD, CQLImpl
from .ddpg_impl import C_ACTOR, C_ALPHA, C_CRITIC, C_DRAW, C_IMITATOR, C_TEMP, _OptimView

M_TEMP_LOSS, M_TEMP, M_ALPHA_LOSS, M_ALPHA, M_CRITIC, M_ACTOR, M_IMITATOR = 0, 1, 2, 3, 4, 5, 6
S_TD, S_ACTOR, S_VAE, S_MMD_ALPHA, S_MMD_ACTOR = 4, 8, 12, 16, 17


class BEARImpl(CQLImpl):
    POLICY_KIND = "normal"

    def __init__(self, *, imitator_learning_rate=3e-4, imitator_hidden=(256, 256), lam=0.75, n_action_samples=100,
                 n_target_samples=10, n_mmd_action_samples=4, mmd_kernel="laplacian", mmd_sigma=20.0, vae_kl_weight=0.5,
                 alpha_threshold=0.05, **kw):
        super().__init__(n_action_samples=n_action_samples, alpha_threshold=alpha_threshold, **kw)
        if mmd_kernel not in ("gaussian", "laplacian"):
            raise ValueError(f"Invalid kernel type: {mmd_kernel}")     # bear_impl.py:248
        self._imitator_learning_rate, self._imitator_hidden = imitator_learning_rate, list(imitator_hidden)
        self._lam, self._n_target, self._n_mmd = lam, int(n_target_samples), int(n_mmd_action_samples)
        self._gaussian, self._mmd_sigma, self._beta = int(mmd_kernel == "gaussian"), float(mmd_sigma), vae_kl_weight

    def build(self) -> None:
        if self.world_size > 1:
            raise NotImplementedError("BEAR: data-parallel exchange is not wired (single GPU only)")
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
        v = super()._checkpoint_views()
        im = self.imitator
        v.update({"_imitator": im,
                  "_imitator_optim": _OptimView(lambda which: im.state_dict(which), self._vae_enc.arena.step,
                                                self._imitator_learning_rate)})
        return v

    def noise_layout(self, B, warmup: bool = False):
        A, n, nt = self._action_size, self._n_mmd, self._n_target
        lay = {"imitator": ("normal", (B, 2 * A))}
        if self._temp_learning_rate > 0:
            lay["temp"] = ("normal", (B, A))
        if self._alpha_learning_rate > 0:
            lay["mmd_lat_alpha"] = ("normal", (n * B, 2 * A))
            lay["mmd_eps_alpha"] = ("normal", (n, B, A))
        lay["target"] = ("normal", (nt, B, A))
        lay["actor"] = ("normal", (B, A))          # unused during warm-up (the reference does not draw it then)
        lay["mmd_lat_actor"] = ("normal", (n * B, 2 * A))
        lay["mmd_eps_actor"] = ("normal", (n, B, A))
        return lay

    # ------------------------------------------------------------------ program pieces
    def _p_imitator(self, db):
        """update_imitator (bear_impl.py:201-213) = ConditionalVAE.compute_error (imitators.py:80-86)."""
        B, O, A, L, st = db.B, db.O, self._action_size, self._lib, self._stream
        Lz = 2 * A
        enc, dec = self._vae_enc, self._vae_dec
        inv_b = 1.0 / B
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
        L.vae_finalize(self.sums_ptr(S_VAE), A, Lz, self._beta, inv_b, self.metric_ptr(M_IMITATOR), st)
        for net in (enc, dec):
            net.adam(self._imitator_learning_rate, st)

    def _p_mmd(self, db, head, tag: str, sum_slot: int, d_head=None):
        """_compute_mmd (bear_impl.py:233-281) on fresh samples; adds the policy-head gradient into d_head when given."""
        B, O, A, n, L, st = db.B, db.O, self._action_size, self._n_mmd, self._lib, self._stream
        Lz = 2 * A
        xd = self.ws("mmd_xd", n * B, O + Lz)
        L.bear_latent_rows(db.ptr("obs"), O, self.noise_view(f"mmd_lat_{tag}", B).data_ptr(), 0.5, xd.data_ptr(), O + Lz,
                           B, n, O, Lz, st)
        raw = self.ws("mmd_raw", 1, n * B, A)
        self._vae_dec.forward("params", xd, O + Lz, n * B, self._vae_dec.ctx("mmd_d", n * B, 1, False), raw, st,
                              head_tanh=False)       # sample_n_without_squash: the decoder's raw `_fc` output
        la = self._log_alpha
        L.bear_mmd(head.data_ptr(), 2 * A, self.noise_view(f"mmd_eps_{tag}", B).data_ptr(), raw.data_ptr(), A,
                   self._gaussian, self._mmd_sigma, MIN_LOGSTD, MAX_LOGSTD, la.ptr("p"), self._alpha_threshold, 1.0 / B,
                   d_head.data_ptr() if d_head is not None else None, 2 * A, self.sums_ptr(sum_slot), B, n, A, st)

    def _p_alpha(self, db, head):
        """update_alpha (bear_impl.py:215-231)."""
        la = self._log_alpha
        self._p_mmd(db, head, "alpha", S_MMD_ALPHA)
        self._lib.bear_alpha_step(self.sums_ptr(S_MMD_ALPHA), la.buf.data_ptr(), self.counter_ptr(C_ALPHA),
                                  self._alpha_learning_rate, 1.0 / db.B, self.metric_ptr(M_ALPHA_LOSS),
                                  self.metric_ptr(M_ALPHA), self._stream)

    def _p_bear_target(self, db, head):
        """compute_target (bear_impl.py:283-303) -> q_tpn[B]."""
        B, O, A, nt, L, st = db.B, db.O, self._action_size, self._n_target, self._lib, self._stream
        R = B * nt
        xt = self.ws("bt_x", R, O + A)
        lp = self.ws("bt_lp", R)
        L.policy_sample_rows(head.data_ptr() + 4 * (B * 2 * A), 2 * A, self.noise_view("target", B).data_ptr(),
                             db.ptr("next_obs"), O, xt.data_ptr(), O + A, None, lp.data_ptr(), B, nt, O, A, MIN_LOGSTD,
                             MAX_LOGSTD, 0, st)
        _, q = self._critic_rows_forward("target", xt, R, "bt_q", train=False)
        q_tpn = self.ws("bt_tpn", B)
        L.bear_target(q.data_ptr(), R, lp.data_ptr(), self._log_temp.ptr("p"), self._lam, q_tpn.data_ptr(), B, nt,
                      self._n_critics, st)
        return q_tpn

    def _p_td(self, db, q_tpn, backward=True):
        B, O, A, E, L, st = db.B, db.O, self._action_size, self._n_critics, self._lib, self._stream
        xc = self.ws("xc", B, O + A)
        L.concat_rows(db.ptr("obs"), O, db.ptr("act"), A, None, 0.0, 0.0, 0.0, xc.data_ptr(), O + A, B, 1, O, A, st)
        acts, q = self._critic_rows_forward("params", xc, B, "cq")
        dq = self.ws("dq", E, B)
        inv_b = 1.0 / B
        L.critic_loss(q.data_ptr(), B, None, B, E, q_tpn.data_ptr(), db.ptr("rew"), db.ptr("term"), db.ptr("nsteps"),
                      self._gamma, None, None, 0, A, None, 0.0, dq.data_ptr(), B, self.sums_ptr(S_TD), None, B, E, inv_b,
                      1, st)
        L.cql_finalize(self.sums_ptr(S_TD), None, inv_b, E, 0.0, 0.0, 0, 0, self.metric_ptr(M_CRITIC), None, st)
        if backward:
            self._q_func.backward(xc, O + A, B, acts, dq, st)
            self._q_func.adam(self._critic_learning_rate, st, tau=self._tau)

    def _p_bear_actor(self, db, acts_p, head, warmup: bool, step=True):
        """warmup_actor / update_actor (bear_impl.py:171-199): [SAC actor loss +] MMD loss, one policy backward."""
        B, O, A, E, L, st = db.B, db.O, self._action_size, self._n_critics, self._lib, self._stream
        inv_b = 1.0 / B
        t = self._log_temp
        dhead = self.ws("pi_dhead", 1, B, 2 * A)
        if warmup:
            L.memset_zero(dhead.data_ptr(), 4 * B * 2 * A, st)
        else:
            xa = self.ws("xa", B, O + A)
            lp = self.ws("a_lp", B)
            eps = self.noise_view("actor", B)
            L.policy_sample_rows(head.data_ptr(), 2 * A, eps.data_ptr(), db.ptr("obs"), O, xa.data_ptr(), O + A, None,
                                 lp.data_ptr(), B, 1, O, A, MIN_LOGSTD, MAX_LOGSTD, 0, st)
            acts_c, q = self._critic_rows_forward("params", xa, B, "aq")
            dqa = self.ws("a_dq", E, B)
            L.sac_actor_loss(q.data_ptr(), B, lp.data_ptr(), t.ptr("p"), dqa.data_ptr(), B, self.sums_ptr(S_ACTOR), B, E,
                             inv_b, st)
            dxa = self.ws("a_dx", E, B, A)
            self._q_func.backward(xa, O + A, B, acts_c, dqa, st, weight_grads=False, dx=dxa, lddx=A, stride_dx=B * A,
                                  dx_col0=O, dx_cols=A)
            L.sac_actor_backward(head.data_ptr(), 2 * A, eps.data_ptr(), dxa.data_ptr(), A, B * A, E, t.ptr("p"),
                                 dhead.data_ptr(), 2 * A, B, A, MIN_LOGSTD, MAX_LOGSTD, inv_b, st)
        self._p_mmd(db, head, "actor", S_MMD_ACTOR, d_head=dhead if step else None)
        L.bear_actor_metric(None if warmup else self.sums_ptr(S_ACTOR), self.sums_ptr(S_MMD_ACTOR),
                            self._log_alpha.ptr("p"), inv_b, self.metric_ptr(M_ACTOR), st)
        if not step:
            return
        self._policy_backward_rows(db, acts_p, dhead, B)
        self._policy.adam(self._actor_learning_rate, st, tau=self._tau)

    # ------------------------------------------------------------------ fused update (BEAR._update)
    def update_fused(self, batch, warmup: bool):
        return self._metrics_dict(self.update_fused_async(batch, warmup))

    def update_fused_async(self, batch, warmup: bool):
        db = self.load_batch(batch, defer=True)
        do_temp, do_alpha = self._temp_learning_rate > 0, self._alpha_learning_rate > 0

        def program():
            ticks = [C_DRAW, C_IMITATOR, C_CRITIC, C_ACTOR] + ([C_TEMP] if do_temp else []) + ([C_ALPHA] if do_alpha else [])
            self._tick(*ticks)
            self.zero_slots()
            self.fill_noise(db.B)
            self._p_imitator(db)
            acts_p, head = self._p_policy(db)          # the policy only changes in the actor step
            if do_temp:
                self._p_temp(db, head)
            if do_alpha:
                self._p_alpha(db, head)
            self._p_td(db, self._p_bear_target(db, head))
            self._p_bear_actor(db, acts_p, head, warmup)

        self.run_program(("bear", db.B, warmup, do_temp, do_alpha, self._noise_injected), program)
        names = [(M_IMITATOR, "imitator_loss")]
        if do_temp:
            names += [(M_TEMP_LOSS, "temp_loss"), (M_TEMP, "temp")]
        if do_alpha:
            names += [(M_ALPHA_LOSS, "alpha_loss"), (M_ALPHA, "alpha")]
        return names + [(M_CRITIC, "critic_loss"), (M_ACTOR, "actor_loss")]

    # ------------------------------------------------------------------ reference hooks (eager)
    def update_imitator(self, batch) -> np.ndarray:
        db = self._begin(batch, C_DRAW, C_IMITATOR)
        self._p_imitator(db)
        return self.read_slots()[M_IMITATOR].copy()

    def update_alpha(self, batch):
        db = self._begin(batch, C_DRAW, C_ALPHA)
        _, head = self._p_policy(db)
        self._p_alpha(db, head)
        v = self.read_slots()
        return v[M_ALPHA_LOSS].copy(), v[M_ALPHA].copy()

    def update_critic(self, batch) -> np.ndarray:
        db = self._begin(batch, C_DRAW, C_CRITIC)
        _, head = self._p_policy(db)
        tau, self._tau = self._tau, None
        try:
            self._p_td(db, self._p_bear_target(db, head))
        finally:
            self._tau = tau
        return self.read_slots()[M_CRITIC].copy()

    def compute_target(self, batch) -> torch.Tensor:
        db = self._begin(batch)
        _, head = self._p_policy(db)
        q_tpn = self._p_bear_target(db, head)
        self.sync()
        return q_tpn.view(-1, 1).clone()

    def _actor_hook(self, batch, warmup: bool, step: bool):
        db = self._begin(batch, *([C_DRAW, C_ACTOR] if step else []))
        acts_p, head = self._p_policy(db)
        tau, self._tau = self._tau, None
        try:
            self._p_bear_actor(db, acts_p, head, warmup, step=step)
        finally:
            self._tau = tau
        return self.read_slots()[M_ACTOR].copy()

    def warmup_actor(self, batch) -> np.ndarray:
        return self._actor_hook(batch, True, True)

    def update_actor(self, batch) -> np.ndarray:
        return self._actor_hook(batch, False, True)

    def compute_actor_loss(self, batch) -> torch.Tensor:
        return torch.tensor(float(self._actor_hook(batch, False, False)))

    # the CQL-specific hooks do not exist on BEAR
    def _compute_conservative_loss(self, *a, **k):
        raise AttributeError("BEARImpl has no conservative loss")

    def compute_critic_loss(self, batch, q_tpn: torch.Tensor) -> torch.Tensor:
        db = self._begin(batch)
        self._p_td(db, q_tpn.to(self._device).reshape(-1).contiguous(), backward=False)
        self.sync()
        return self._slots[M_CRITIC].clone()

    def _predict_best_action(self, obs: torch.Tensor) -> torch.Tensor:
        """bear_impl.py:305-329: n_action_samples policy samples per observation (onnx_safe_sample_n: tanh(mu + std *
        randn)), the one critic 0 scores highest."""
        n, O, A, N, L, st = obs.shape[0], obs.shape[1], self._action_size, self._n_action_samples, self._lib, self._stream
        head = self._policy_head(obs, head_tanh=False)
        with torch.cuda.stream(self._stream_obj):
            eps = torch.randn(N, n, A, device=self._device)
        x = self.ws("e_x", n * N, O + A)
        L.policy_sample_rows(head.data_ptr(), 2 * A, eps.data_ptr(), obs.data_ptr(), O, x.data_ptr(), O + A, None, None,
                             n, N, O, A, MIN_LOGSTD, MAX_LOGSTD, 0, st)
        _, q = self._critic_rows_forward("params", x, n * N, "e_q", members=1, train=False)
        with torch.cuda.stream(self._stream_obj):
            index = q[0, :n * N].view(n, N).argmax(dim=1)
            return x[:, O:O + A].reshape(n, N, A)[torch.arange(n, device=self._device), index].clone()
