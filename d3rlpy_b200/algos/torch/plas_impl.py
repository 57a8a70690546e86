"""PLAS on B200: mirrors PLASImpl (d3rlpy/algos/torch/plas_impl.py:25-168) over the BCQ building blocks — the same
conditional VAE (imitators.py:13-118) and critic step, with a DeterministicPolicy that acts in the VAE's latent space:
action = decode(s, 2 * pi(s)).  One update (PLAS._update, algos/plas.py:189-206):

  warm-up (grad_step < warmup_steps): the VAE step alone
  afterwards: critic step against decode(s', 2 pi'(s')) through the target critics, lam-mix of the members' min / max
              (ensemble_q_function.py:9-24 "mix"); every `update_actor_interval` steps the actor step
              -mean Q_0(s, decode(s, 2 pi(s))) — gradient through critic 0 and the (frozen) decoder into the policy —
              then the actor-target and critic-target soft syncs."""
from __future__ import annotations

import numpy as np
import torch

from ...nets import DenseNet
from .bcq_impl import M_ACTOR, M_CRITIC, M_IMITATOR, S_ACT, BCQImpl
from .ddpg_impl import C_ACTOR, C_CRITIC, C_DRAW, C_IMITATOR


class PLASImpl(BCQImpl):
    POLICY_KIND = "plas"

    def __init__(self, **kw):
        kw.setdefault("n_action_samples", 1)
        super().__init__(**kw)

    def _build_actor(self) -> None:
        O, A = self._observation_shape[0], self._action_size
        self._policy = DenseNet(O, self._actor_hidden, [("_fc", 2 * A)], 1, self._device, trunk_prefix="_encoder.",
                                with_target=True, seed_gen=self._gen, precision=self._precision)

    def noise_layout(self, B):
        return {"imitator": ("normal", (B, 2 * self._action_size))}

    # ------------------------------------------------------------------ program pieces
    def _decode(self, which: str, db, field: str, tag: str, train: bool):
        """decode(s, 2 * pi(s)) -> (policy ctx, z = tanh latent [1,B,2A], decoder rows, decoder ctx, action [1,B,A])."""
        B, O, A, L, st = db.B, db.O, self._action_size, self._lib, self._stream
        Lz = 2 * A
        cp = self._policy.ctx(f"{tag}_p", B, 1, train)
        z = self.ws(f"{tag}_z", 1, B, Lz)
        self._policy.forward(which, db.ptr(field), O, B, cp, z, st, head_tanh=True)
        xd = self.ws(f"{tag}_xd", B, O + Lz)
        L.scaled_concat_rows(db.ptr(field), O, z.data_ptr(), Lz, 2.0, xd.data_ptr(), O + Lz, B, O, Lz, st)
        cd = self._vae_dec.ctx(f"{tag}_d", B, 1, train)
        act = self.ws(f"{tag}_a", 1, B, A)
        self._vae_dec.forward("params", xd, O + Lz, B, cd, act, st, head_tanh=True)
        return cp, z, xd, cd, act

    def _p_target(self, db):
        """compute_target (plas_impl.py:159-168): mix of the target members at decode(s', 2 pi'(s')) -> q_tpn[B]."""
        B, O, A, L, st = db.B, db.O, self._action_size, self._lib, self._stream
        _, _, _, _, act = self._decode("target", db, "next_obs", "t", False)
        xq = self.ws("t_xq", B, O + A)
        L.concat_rows(db.ptr("next_obs"), O, act.data_ptr(), A, None, 0.0, 0.0, 0.0, xq.data_ptr(), O + A, B, 1, O, A, st)
        _, q = self._critic_rows_forward("target", xq, B, "t_q", train=False)
        q_tpn = self.ws("q_tpn", B)
        L.ensemble_reduce(q.data_ptr(), B, B, self._n_critics, 3, self._lam, q_tpn.data_ptr(), st)
        return q_tpn

    def _p_actor(self, db, sync_target=True, step=True):
        """compute_actor_loss (plas_impl.py:138-147) [+ backward + Adam + soft syncs]."""
        B, O, A, L, st = db.B, db.O, self._action_size, self._lib, self._stream
        Lz = 2 * A
        cp, z, xd, cd, act = self._decode("params", db, "obs", "a", True)
        xq = self.ws("a_xq", B, O + A)
        L.concat_rows(db.ptr("obs"), O, act.data_ptr(), A, None, 0.0, 0.0, 0.0, xq.data_ptr(), O + A, B, 1, O, A, st)
        cc, q0 = self._critic_rows_forward("params", xq, B, "aq", members=1)
        dq = self.ws("a_dq", 1, B)
        inv_b = 1.0 / B
        L.neg_mean_seed(q0.data_ptr(), dq.data_ptr(), self.sums_ptr(S_ACT), B, inv_b, st)
        L.copy_d2d(self.metric_ptr(M_ACTOR), self.sums_ptr(S_ACT), 4, st)
        if not step:
            return
        da = self.ws("a_da", B, A)
        self._q_func.backward(xq, O + A, B, cc, dq, st, weight_grads=False, dx=da, lddx=A, stride_dx=B * A, dx_col0=O,
                              dx_cols=A)
        dpre = self.ws("a_dpre", 1, B, A)                      # through the decoder's tanh
        L.tanh_backward(da.data_ptr(), A, act.data_ptr(), A, 1.0, dpre.data_ptr(), A, B, A, st)
        dlat = self.ws("a_dlat", B, Lz)                        # decoder data gradient w.r.t. the latent columns
        self._vae_dec.backward(xd, O + Lz, B, cd, dpre, st, weight_grads=False, dx=dlat, lddx=Lz, stride_dx=B * Lz,
                               dx_col0=O, dx_cols=Lz)
        dz = self.ws("a_dz", 1, B, Lz)                         # latent = 2 * tanh(fc): factor 2 and the policy's tanh
        L.tanh_backward(dlat.data_ptr(), Lz, z.data_ptr(), Lz, 2.0, dz.data_ptr(), Lz, B, Lz, st)
        self._policy.backward(db.ptr("obs"), O, B, cp, dz, st)
        self._policy.adam(self._actor_learning_rate, st, tau=self._tau if sync_target else None)

    # ------------------------------------------------------------------ fused update (PLAS._update)
    def update_fused(self, batch, warmup: bool, actor_step: bool):
        return self._metrics_dict(self.update_fused_async(batch, warmup, actor_step))

    def update_fused_async(self, batch, warmup: bool, actor_step: bool):
        db = self.load_batch(batch, defer=True)
        actor_step = actor_step and not warmup

        def program():
            if warmup:
                self._tick(C_DRAW, C_IMITATOR)
                self.zero_slots()
                self.fill_noise(db.B)
                self._p_imitator(db)
                return
            self._tick(C_CRITIC, *([C_ACTOR] if actor_step else []))
            self.zero_slots()
            q_tpn = self._p_target(db)
            # reference order: critic step, actor step, actor-target sync, critic-target sync (plas.py:197-204); the
            # critic target depends only on the critic parameters, so its soft sync rides in the critic Adam pass
            self._p_critic(db, q_tpn, sync_target=actor_step)
            if actor_step:
                self._p_actor(db)

        self.run_program(("plas", db.B, warmup, actor_step, self._noise_injected), program)
        if warmup:
            return [(M_IMITATOR, "imitator_loss")]
        return [(M_CRITIC, "critic_loss")] + ([(M_ACTOR, "actor_loss")] if actor_step else [])

    # ------------------------------------------------------------------ reference hooks / evaluation
    def update_critic(self, batch) -> np.ndarray:
        db = self.load_batch(batch)
        self._tick(C_CRITIC)
        self.zero_slots()
        self._p_critic(db, self._p_target(db))
        return self.read_slots()[M_CRITIC].copy()

    def update_actor(self, batch) -> np.ndarray:
        db = self.load_batch(batch)
        self._tick(C_ACTOR)
        self.zero_slots()
        self._p_actor(db, sync_target=False)
        return self.read_slots()[M_ACTOR].copy()

    def compute_actor_loss(self, batch) -> torch.Tensor:
        db = self.load_batch(batch)
        self.zero_slots()
        self._p_actor(db, step=False)
        self.sync()
        return self._slots[M_ACTOR].clone()

    def _predict_best_action(self, obs: torch.Tensor, latent=None) -> torch.Tensor:
        """plas_impl.py:149-151: decode(x, 2 * policy(x))."""
        n, O, A, L, st = obs.shape[0], obs.shape[1], self._action_size, self._lib, self._stream
        Lz = 2 * A
        z = self.ws("e_z", 1, n, Lz)
        self._policy.forward("params", obs, O, n, self._policy.ctx("e_p", n, 1, False), z, st, head_tanh=True)
        xd = self.ws("e_xd", n, O + Lz)
        L.scaled_concat_rows(obs.data_ptr(), O, z.data_ptr(), Lz, 2.0, xd.data_ptr(), O + Lz, n, O, Lz, st)
        act = self.ws("e_a", 1, n, A)
        self._vae_dec.forward("params", xd, O + Lz, n, self._vae_dec.ctx("e_d", n, 1, False), act, st, head_tanh=True)
        return act[0].clone()

    def sample_action(self, x) -> np.ndarray:
        return self.predict_best_action(x)   # plas_impl.py:153-154
