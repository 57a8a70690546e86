"""CQL (SAC-based) on B200: mirrors CQLImpl / SACImpl
(d3rlpy/algos/torch/cql_impl.py:21-243, sac_impl.py:33-162).

One update = ONE policy-trunk forward over [obs; next_obs] (the policy does not change until
update_actor, so temp/alpha/critic/actor steps share it), two critic passes over the
B(1+3N) importance-sampling rows (alpha step: forward only; critic step: forward + backward), a
B-row target pass and a B-row actor pass with data-gradient only — the reference's wasted critic
backward in update_alpha and the unused critic weight-gradients in update_actor are not computed
(SURVEY.md §8d "req").
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from ...nets import DenseNet
from .ddpg_impl import C_ACTOR, C_ALPHA, C_CRITIC, C_DRAW, C_TEMP, DDPGBaseImpl

M_TEMP_LOSS, M_TEMP, M_ALPHA_LOSS, M_ALPHA, M_CRITIC, M_ACTOR = 0, 1, 2, 3, 4, 5
S_ALPHA, S_CRITIC, S_ACTOR = 0, 4, 8
MIN_LOGSTD, MAX_LOGSTD = -20.0, 2.0  # create_squashed_normal_policy defaults (models/builders.py:105-121)


class _Scalar:
    """A (1,1) nn.Parameter-like scalar (d3rlpy/models/torch/parameters.py:5-21) with Adam state."""

    def __init__(self, value: float, device, step_view):
        self.buf = torch.zeros(16, dtype=torch.float32, device=device)  # p, g, m, v at 16B-aligned slots
        self.buf[0] = value
        self.step = step_view

    def ptr(self, which: str) -> int:
        return self.buf.data_ptr() + 16 * {"p": 0, "g": 1, "m": 2, "v": 3}[which]

    @property
    def data(self) -> torch.Tensor:
        return self.buf[0:1].view(1, 1)

    def state_dict(self):
        return {"_parameter": self.data}

    def load_state_dict(self, sd):
        self.buf[0] = float(sd["_parameter"].reshape(-1)[0])

    def optim_view(self, lr: float):
        """Adam state of the scalar in torch.optim layout (one (1,1) parameter)."""
        from collections import OrderedDict

        def sd_of(which):
            i = {"params": 0, "exp_avg": 8, "exp_avg_sq": 12}[which]
            return OrderedDict([("_parameter", self.buf[i:i + 1].view(1, 1))])

        from .ddpg_impl import _OptimView
        return _OptimView(sd_of, self.step, lr)


class CQLImpl(DDPGBaseImpl):
    POLICY_KIND = "normal"
    def __init__(self, *, temp_learning_rate=1e-4, alpha_learning_rate=1e-4, initial_temperature=1.0,
                 initial_alpha=1.0, alpha_threshold=10.0, conservative_weight=5.0, n_action_samples=10,
                 soft_q_backup=False, **kw):
        super().__init__(**kw)
        self._temp_learning_rate, self._alpha_learning_rate = temp_learning_rate, alpha_learning_rate
        self._initial_temperature, self._initial_alpha = initial_temperature, initial_alpha
        self._alpha_threshold, self._conservative_weight = alpha_threshold, conservative_weight
        self._n_action_samples, self._soft_q_backup = n_action_samples, soft_q_backup

    def build(self) -> None:
        super().build()
        self._log_temp = _Scalar(math.log(self._initial_temperature), self._device, self._counters[C_TEMP:C_TEMP + 1])
        self._log_alpha = _Scalar(math.log(self._initial_alpha), self._device, self._counters[C_ALPHA:C_ALPHA + 1])

    def _build_actor(self) -> None:
        O, A = self._observation_shape[0], self._action_size
        self._policy = DenseNet(O, self._actor_hidden, [("_mu", A), ("_logstd", A)], 1, self._device,
                                trunk_prefix="_encoder.", with_target=True, seed_gen=self._gen,
                                precision=self._precision)

    def noise_layout(self, B):
        """Reference draw order per update."""
        A, N = self._action_size, self._n_action_samples
        lay = {}
        if self._temp_learning_rate > 0:
            lay["temp"] = ("normal", (B, A))
        if self._alpha_learning_rate > 0:
            lay["alpha_t"] = ("normal", (N, B, A))
            lay["alpha_tp1"] = ("normal", (N, B, A))
            lay["alpha_rand"] = ("uniform", (B * N, A))
        if self._soft_q_backup:
            lay["soft"] = ("normal", (B, A))
        lay["critic_t"] = ("normal", (N, B, A))
        lay["critic_tp1"] = ("normal", (N, B, A))
        lay["critic_rand"] = ("uniform", (B * N, A))
        lay["actor"] = ("normal", (B, A))
        return lay

    # ------------------------------------------------------------------ program pieces
    def _p_policy(self, db, converted: bool = False):
        """Policy trunk + (mu|logstd) head on [obs; next_obs] (contiguous in the device batch).  converted: the
        update prologue already wrote the bf16 operand rows into the workspace."""
        B, O, A = db.B, db.O, self._action_size
        acts = self._policy.ctx("pi", 2 * B, 1, True)
        head = self.ws("pi_head", 1, 2 * B, 2 * A)
        assert db.off["next_obs"] == db.off["obs"] + B * O, "obs/next_obs must be contiguous"
        self._policy.forward("params", db.ptr("obs"), O, 2 * B, acts, head, self._stream,
                             x_bf16=(acts.xb.data_ptr(), acts.ldk0) if converted else None)
        return acts, head

    def _p_temp(self, db, head):
        """update_temp (sac_impl.py:123-146)."""
        B, A, L, st = db.B, self._action_size, self._lib, self._stream
        lp = self.ws("temp_lp", B)
        L.policy_sample_rows(head.data_ptr(), 2 * A, self.noise_view("temp", B).data_ptr(), None, 0, None, 0, None,
                             lp.data_ptr(), B, 1, 0, A, MIN_LOGSTD, MAX_LOGSTD, 0, st)
        self._allreduce_pre_scalar(lp, B)
        t = self._log_temp
        L.sac_temp_loss(lp.data_ptr(), t.ptr("p"), B, A, 1.0 / (B * self.world_size), self.metric_ptr(M_TEMP_LOSS), t.ptr("g"), 0, st)
        self._allreduce_scalar_grad(t, M_TEMP_LOSS)
        L.scalar_adam(t.ptr("p"), t.ptr("g"), t.ptr("m"), t.ptr("v"), self.counter_ptr(C_TEMP),
                      self._temp_learning_rate, 0.9, 0.999, 1e-8, self.metric_ptr(M_TEMP), st)

    def _p_rows(self, db, head, tag):
        """Critic input rows [data | pi(s_t) | pi(s_t+1) | random] at obs_t (cql_impl.py:143-204)."""
        B, O, A, N, L, st = db.B, db.O, self._action_size, self._n_action_samples, self._lib, self._stream
        R = B * (1 + 3 * N)
        ld = O + A
        x = self.ws("x_is", R, ld)
        lp = self.ws("lp_is", 2, max(B * N, 1))
        xp = x.data_ptr()
        L.concat_rows(db.ptr("obs"), O, db.ptr("act"), A, None, 0.0, 0.0, 0.0, xp, ld, B, 1, O, A, st)
        if N == 0:  # plain SAC: the data rows are the whole critic input
            return x, lp, R
        L.policy_sample_rows(head.data_ptr(), 2 * A, self.noise_view(f"{tag}_t", B).data_ptr(), db.ptr("obs"), O,
                             xp + 4 * ld * B, ld, None, lp.data_ptr(), B, N, O, A, MIN_LOGSTD, MAX_LOGSTD, 0, st)
        L.policy_sample_rows(head.data_ptr() + 4 * (B * 2 * A), 2 * A, self.noise_view(f"{tag}_tp1", B).data_ptr(),
                             db.ptr("obs"), O, xp + 4 * ld * (B + B * N), ld, None, lp.data_ptr() + 4 * B * N, B, N, O,
                             A, MIN_LOGSTD, MAX_LOGSTD, 0, st)
        L.concat_rows(db.ptr("obs"), O, self.noise_view(f"{tag}_rand", B).data_ptr(), A, None, 0.0, 0.0, 0.0,
                      xp + 4 * ld * (B + 2 * B * N), ld, B, N, O, A, st)
        return x, lp, R

    def _p_alpha(self, db, head):
        """update_alpha (cql_impl.py:119-141): forward only — the alpha gradient needs no critic backward."""
        B, A, N, L, st, E = db.B, self._action_size, self._n_action_samples, self._lib, self._stream, self._n_critics
        x, lp, R = self._p_rows(db, head, "alpha")
        _, q = self._critic_rows_forward("params", x, R, "is")
        inv_b = 1.0 / (B * self.world_size)
        la = self._log_alpha
        L.critic_loss(q.data_ptr(), R, None, 0, 0, None, None, None, None, self._gamma, lp.data_ptr(),
                      lp.data_ptr() + 4 * B * N, N, A, la.ptr("p"), self._conservative_weight, None, 0,
                      self.sums_ptr(S_ALPHA), None, B, E, inv_b, 0, st)
        self._allreduce(self._slots[32 + S_ALPHA:32 + S_ALPHA + 3])
        L.cql_finalize(self.sums_ptr(S_ALPHA), la.ptr("p"), inv_b, E, self._conservative_weight,
                       self._alpha_threshold, 1, 1, self.metric_ptr(M_ALPHA_LOSS), la.ptr("g"), st)
        L.scalar_adam(la.ptr("p"), la.ptr("g"), la.ptr("m"), la.ptr("v"), self.counter_ptr(C_ALPHA),
                      self._alpha_learning_rate, 0.9, 0.999, 1e-8, self.metric_ptr(M_ALPHA), st)

    def _p_target(self, db, head):
        """compute_target (cql_impl.py:225-243): deterministic backup tanh(mu(s')) through the target critics
        (returns q_t[E,B]; the min over members is taken inside critic_loss), or with soft_q_backup the SAC
        target min_e Q' - exp(log_temp) logp with a sampled action (sac_impl.py:148-162; returns q_tpn[B])."""
        B, O, A, L, st = db.B, db.O, self._action_size, self._lib, self._stream
        xt = self.ws("xt", B, O + A)
        if not self._soft_q_backup:
            L.policy_sample_rows(head.data_ptr() + 4 * (B * 2 * A), 2 * A, None, db.ptr("next_obs"), O, xt.data_ptr(),
                                 O + A, None, None, B, 1, O, A, MIN_LOGSTD, MAX_LOGSTD, 1, st)
            _, q_t = self._critic_rows_forward("target", xt, B, "tq", train=False)
            return q_t, None
        lp = self.ws("soft_lp", B)
        L.policy_sample_rows(head.data_ptr() + 4 * (B * 2 * A), 2 * A, self.noise_view("soft", B).data_ptr(),
                             db.ptr("next_obs"), O, xt.data_ptr(), O + A, None, lp.data_ptr(), B, 1, O, A, MIN_LOGSTD,
                             MAX_LOGSTD, 0, st)
        _, q_t = self._critic_rows_forward("target", xt, B, "tq", train=False)
        q_tpn = self.ws("soft_tpn", B)
        L.sac_soft_backup(q_t.data_ptr(), B, self._n_critics, lp.data_ptr(), self._log_temp.ptr("p"),
                          q_tpn.data_ptr(), B, st)
        return None, q_tpn

    def _p_critic(self, db, head, q_t=None, q_tpn=None, backward=True, sync_target=True, conservative=True,
                  td=True):
        """compute_critic_loss (cql_impl.py:110-117) [+ backward + Adam (ddpg_impl.py:138-152)]."""
        B, A, N, L, st, E = db.B, self._action_size, self._n_action_samples, self._lib, self._stream, self._n_critics
        x, lp, R = self._p_rows(db, head, "critic")
        acts, q = self._critic_rows_forward("params", x, R, "is")
        dq = self.ws("is_dq", E, R)
        inv_b = 1.0 / (B * self.world_size)
        la = self._log_alpha
        L.critic_loss(q.data_ptr(), R, q_t.data_ptr() if q_t is not None else None, B, E,
                      q_tpn.data_ptr() if q_tpn is not None else None, db.ptr("rew"), db.ptr("term"),
                      db.ptr("nsteps"), self._gamma, lp.data_ptr(), lp.data_ptr() + 4 * B * N, N if conservative else 0,
                      A, la.ptr("p"), self._conservative_weight, dq.data_ptr(), R, self.sums_ptr(S_CRITIC), None, B, E,
                      inv_b, 1 if td else 0, st)
        self._allreduce(self._slots[32 + S_CRITIC:32 + S_CRITIC + 3])
        L.cql_finalize(self.sums_ptr(S_CRITIC), la.ptr("p"), inv_b, E, self._conservative_weight,
                       self._alpha_threshold, 0, 1 if conservative else 0, self.metric_ptr(M_CRITIC), None, st)
        if backward:
            self._q_func.backward(x, self._q_func.in_dim, R, acts, dq, st)
            self._allreduce(self._q_func.arena.grads)
            self._q_func.adam(self._critic_learning_rate, st, tau=self._tau if sync_target else None)

    def _p_actor(self, db, acts_p, head, sync_target=True, step=True):
        """compute_actor_loss (sac_impl.py:114-121) + backward + Adam (ddpg_impl.py:167-183); step=False stops after
        the loss value."""
        B, O, A, L, st, E = db.B, db.O, self._action_size, self._lib, self._stream, self._n_critics
        xa = self.ws("xa", B, O + A)
        lp = self.ws("a_lp", B)
        eps = self.noise_view("actor", B)
        L.policy_sample_rows(head.data_ptr(), 2 * A, eps.data_ptr(), db.ptr("obs"), O, xa.data_ptr(), O + A, None,
                             lp.data_ptr(), B, 1, O, A, MIN_LOGSTD, MAX_LOGSTD, 0, st)
        acts_c, q = self._critic_rows_forward("params", xa, B, "aq")
        dq = self.ws("a_dq", E, B)
        inv_b = 1.0 / (B * self.world_size)
        t = self._log_temp
        L.sac_actor_loss(q.data_ptr(), B, lp.data_ptr(), t.ptr("p"), dq.data_ptr(), B, self.sums_ptr(S_ACTOR), B, E,
                         inv_b, st)
        self._allreduce(self._slots[32 + S_ACTOR:32 + S_ACTOR + 1])
        L.copy_d2d(self.metric_ptr(M_ACTOR), self.sums_ptr(S_ACTOR), 4, st)
        if not step:
            return
        dxa = self.ws("a_dx", E, B, A)
        self._q_func.backward(xa, O + A, B, acts_c, dq, st, weight_grads=False, dx=dxa, lddx=A, stride_dx=B * A,
                              dx_col0=O, dx_cols=A)
        dhead = self.ws("pi_dhead", 1, B, 2 * A)
        L.sac_actor_backward(head.data_ptr(), 2 * A, eps.data_ptr(), dxa.data_ptr(), A, B * A, E, t.ptr("p"),
                             dhead.data_ptr(), 2 * A, B, A, MIN_LOGSTD, MAX_LOGSTD, inv_b, st)
        # policy backward over the first B rows (obs_t) of the shared [obs; next_obs] forward
        self._policy_backward_rows(db, acts_p, dhead, B)
        self._allreduce(self._policy.arena.grads)
        self._policy.adam(self._actor_learning_rate, st, tau=self._tau if sync_target else None)

    def _policy_backward_rows(self, db, acts_p, dhead, B):
        """acts_p hold 2B rows ([obs; next_obs]); E == 1 so the first B rows are a contiguous prefix."""
        self._policy.backward(db.ptr("obs"), db.O, B, acts_p, dhead, self._stream)

    def _allreduce(self, t):
        if self.world_size > 1:
            from ...parallel import allreduce_sum

            allreduce_sum(t, self._stream_obj)

    def _allreduce_pre_scalar(self, lp, B):
        pass

    def _allreduce_scalar_grad(self, scalar, metric_slot):
        if self.world_size > 1:
            from ...parallel import allreduce_sum

            allreduce_sum(scalar.buf[4:5], self._stream_obj)
            allreduce_sum(self._slots[metric_slot:metric_slot + 1], self._stream_obj)

    # ------------------------------------------------------------------ single-GPU tensor-core program
    fused_glue = True  # collapse the glue between the GEMM launches (csrc/cql_fused.cu); False = generic path

    def _side_stream(self) -> int:
        if getattr(self, "_side_obj", None) is None:
            self._side_obj = torch.cuda.Stream(device=self._device)
        return self._side_obj.cuda_stream

    def _alpha_stream(self) -> int:
        if getattr(self, "_alpha_obj", None) is None:
            self._alpha_obj = torch.cuda.Stream(device=self._device, priority=-1)   # high priority: its blocks go first
        return self._alpha_obj.cuda_stream

    def _program_fused(self, db, do_temp, do_alpha):
        """The same update as `program` in update_fused_async with ~half the launches: one row-assembly
        kernel, the alpha-step and critic-step critic forwards in ONE launch, loss + scalar tails fused."""
        import ctypes

        B, O, A, N, E = db.B, db.O, self._action_size, self._n_action_samples, self._n_critics
        L, st = self._lib, self._stream
        f32 = self._precision == "fp32"   # fp32 mode: fp32 operand rows, one GEMM launch per layer (3xTF32 / SIMT)
        R = B * (1 + 3 * N)
        G = 2 if do_alpha else 1
        ld = (O + A + 3) // 4 * 4 if f32 else (O + A + 7) // 8 * 8
        esz = 4 if f32 else 2
        X = self.ws("xf_rows", G * R + 2 * B, ld, dtype=torch.float32 if f32 else torch.bfloat16)
        xrows = lambda row0: X.data_ptr() + esz * row0 * ld

        def q_forward(which, row0, rows, ctx, out, stream, save_rows=0):
            if f32:
                q_net.forward(which, xrows(row0), ld, rows, ctx, out, stream)
            else:
                q_net.forward(which, None, 0, rows, ctx, out, stream, x_bf16=(xrows(row0), ld), save_rows=save_rows)

        def q_backward(row0, rows, ctx, dq_, stream, **kw):
            if f32:
                q_net.backward(xrows(row0), ld, rows, ctx, dq_, stream, **kw)
            else:
                q_net.backward(None, 0, rows, ctx, dq_, stream, **kw)
        lp = self.ws("xf_lp", 4, max(B * N, 1))
        lpm = self.ws("xf_lpm", 3, B)  # soft-backup, actor, temp log-probs
        # per loss kernel: block counter + per-block partial sums (fixed-order final sums, csrc/cql_fused.cu)
        done = self.ws("xf_done", 4, 4 + 3 * ((B * E + 7) // 8) + 4, dtype=torch.int32)
        dp = self.world_size > 1  # sharded minibatch: sums / gradients are all-reduced between the partial kernels
        inv_b = 1.0 / (B * self.world_size)
        mask = 0
        for c in [C_DRAW, C_CRITIC, C_ACTOR] + ([C_TEMP] if do_temp else []) + ([C_ALPHA] if do_alpha else []):
            mask |= 1 << c
        px = getattr(self, "_px", None) if dp else None
        one_prologue = os.environ.get("D3B_PROLOGUE", "1") != "0"   # 0: the three separate launches (A/B timing)
        if one_prologue:
            # ONE prologue launch: counters, slot zeroing, the update's noise, bf16 operand rows of the policy input
            _, n_norm, n_uni, _ = self._noise_plan(B)
            draw = not self._noise_injected and n_norm + n_uni > 0
            pre = self._policy.ctx("pi", 2 * B, 1, True) if not f32 else None
            L.update_prologue(self._counters.data_ptr(), self.N_COUNTERS, mask, C_DRAW, self._slots.data_ptr(), 64,
                              self._noise_arena(B).data_ptr() if draw else None, n_norm if draw else 0,
                              n_uni if draw else 0,
                              (self._seed + 0x9E3779B97F4A7C15 * self.rank) & 0xFFFFFFFFFFFFFFFF,   # per-rank Philox key
                              db.ptr("obs") if pre is not None else None, O, 2 * B, O,
                              pre.xb.data_ptr() if pre is not None else None, pre.ldk0 if pre is not None else 0,
                              done[3].data_ptr(), st)
        else:
            L.begin_step(self._counters.data_ptr(), self.N_COUNTERS, mask, self._slots.data_ptr(), 64, st)
        if px is not None:
            # peers have finished reading last update's gradients -> zero them for this update's RED accumulation
            fq, fp = self._q_func._peer[1], self._policy._peer[1]
            L.peer_wait_zero(px.flags_ptrs, px.world, px.rank, fq + 1, self.counter_ptr(C_DRAW),
                             self._q_func.arena.grads.data_ptr(), self._q_func.arena.size, fp + 1,
                             self._policy.arena.grads.data_ptr(), self._policy.arena.size, st)
        if not one_prologue:
            self.fill_noise(B)
        acts_p, head = self._p_policy(db, converted=one_prologue and not f32)
        nv = lambda name: self.noise_view(name, B).data_ptr()
        soft = self._soft_q_backup
        ptrs = [nv("critic_t"), nv("critic_tp1"), nv("critic_rand"), lp[0].data_ptr(), lp[1].data_ptr()] if N > 0 \
            else [None] * 5  # N == 0: plain SAC (algos/torch/sac_impl.py), data rows only
        ptrs += [nv("alpha_t"), nv("alpha_tp1"), nv("alpha_rand"), lp[2].data_ptr(), lp[3].data_ptr()] if do_alpha \
            else [None] * 5
        ptrs += [nv("soft") if soft else None, lpm[0].data_ptr() if soft else None, nv("actor"), lpm[1].data_ptr(),
                 nv("temp") if do_temp else None, lpm[2].data_ptr() if do_temp else None]
        t_row0, a_row0 = G * R, G * R + B
        (L.cql_rows_f32 if f32 else L.cql_rows)(head.data_ptr(), db.ptr("obs"), db.ptr("next_obs"), db.ptr("act"), B, N,
                                                O, A, MIN_LOGSTD, MAX_LOGSTD, X.data_ptr(), ld, G,
                                                (ctypes.c_void_p * 16)(*ptrs),
                                                (ctypes.c_int64 * 4)(0, R, t_row0, a_row0), st)
        la, lt = self._log_alpha, self._log_temp
        q_net = self._q_func
        # ---- side branch (independent of the importance-sampling pass): temperature step, target critics
        side = self._side_stream()
        # data parallel: per-rank partial loss (== gradient of log_temp) -> all-reduce -> Adam.  Collectives stay on
        # the main stream (graph branches could reorder them across ranks).  Unless the soft backup needs the new
        # temperature for the target, the exchange is merged with the alpha step's (one rendezvous instead of two):
        # the temperature partial sum lands in the free fourth float of the alpha sums.
        temp_merged = do_temp and dp and do_alpha and not soft
        temp_g = self.sums_ptr(S_ALPHA + 3) if temp_merged else lt.ptr("g")

        def temp_adam():
            L.copy_d2d(self.metric_ptr(M_TEMP_LOSS), temp_g, 4, st)
            L.scalar_adam(lt.ptr("p"), temp_g, lt.ptr("m"), lt.ptr("v"), self.counter_ptr(C_TEMP),
                          self._temp_learning_rate, 0.9, 0.999, 1e-8, self.metric_ptr(M_TEMP), st)

        if do_temp and dp:
            L.sac_temp_loss(lpm[2].data_ptr(), lt.ptr("p"), B, A, inv_b, self.metric_ptr(M_TEMP_LOSS), temp_g, 0, st)
            if not temp_merged:
                self._small_allreduce(px, lt.buf[4:5], 0)
                temp_adam()
        L.stream_fork(st, side)
        if do_temp and not dp:
            L.sac_temp_step(lpm[2].data_ptr(), lt.buf.data_ptr(), self.counter_ptr(C_TEMP), B, A, inv_b,
                            self._temp_learning_rate, self.metric_ptr(M_TEMP_LOSS), self.metric_ptr(M_TEMP), side)
        ctx_t = q_net.ctx("tq", B, E, False)
        q_t = self.ws("tq_q", E, B)
        q_forward("target", t_row0, B, ctx_t, q_t, side)
        q_tpn = None
        if soft:
            q_tpn = self.ws("soft_tpn", B)
            L.sac_soft_backup(q_t.data_ptr(), B, E, lpm[0].data_ptr(), lt.ptr("p"), q_tpn.data_ptr(), B, side)
        # ---- main branch: critic-step rows [0,R) (activations saved) and alpha-step rows [R,2R) (forward only)
        alpha_branch = do_alpha and (not dp or px is not None) and os.environ.get("D3B_ALPHA_BRANCH", "1") != "0"
        if alpha_branch:
            # The alpha-step rows get their own forward launch on a high-priority branch, followed by the alpha loss
            # step: its blocks are placed first, so log_alpha is already updated when the critic-step rows (main
            # stream, second wave of the same 2 x 124 units) finish — the alpha loss leaves the critical path.
            side_a = self._alpha_stream()
            L.stream_fork(st, side_a)
            q_al = self.ws("is_al_q", E, R)
            q_forward("params", R, R, q_net.ctx("is_al", R, E, False), q_al, side_a)
            if not dp:
                L.cql_loss_step(q_al.data_ptr(), R, None, 0, 0, None, None, None, None, self._gamma,
                                lp[2].data_ptr(), lp[3].data_ptr(), N, A, la.buf.data_ptr(), self._conservative_weight,
                                self._alpha_threshold, None, 0, self.sums_ptr(S_ALPHA), done[0].data_ptr(), B, E, inv_b,
                                1, self.counter_ptr(C_ALPHA), self._alpha_learning_rate,
                                self.metric_ptr(M_ALPHA_LOSS), self.metric_ptr(M_ALPHA), side_a)
            else:
                # sharded: partial sums, then exchange + temperature step + alpha step in one launch — the rendezvous
                # of the scalar steps runs on this branch too (its flags / exchange slots are its own channel)
                L.critic_loss(q_al.data_ptr(), R, None, 0, 0, None, None, None, None, self._gamma, lp[2].data_ptr(),
                              lp[3].data_ptr(), N, A, la.ptr("p"), self._conservative_weight, None, 0,
                              self.sums_ptr(S_ALPHA), None, B, E, inv_b, 0, side_a)
                L.dp_scalar_steps(self.sums_ptr(S_ALPHA), px.xchg_ptrs, px.flags_ptrs, px.world, px.rank, 1,
                                  self.counter_ptr(C_DRAW), lt.buf.data_ptr() if temp_merged else None,
                                  self.counter_ptr(C_TEMP), self._temp_learning_rate, self.metric_ptr(M_TEMP_LOSS),
                                  self.metric_ptr(M_TEMP), la.buf.data_ptr(), self.counter_ptr(C_ALPHA),
                                  self._alpha_learning_rate, inv_b / E, self._conservative_weight,
                                  self._alpha_threshold, self.metric_ptr(M_ALPHA_LOSS), self.metric_ptr(M_ALPHA), side_a)
            GR = R
        else:
            GR = G * R
        ctx = q_net.ctx("is2", GR, E, True)
        q = self.ws("is2_q", E, GR)
        q_forward("params", 0, GR, ctx, q, st, save_rows=R)
        if alpha_branch:
            L.stream_join(st, side_a)
        elif do_alpha and not dp:
            L.cql_loss_step(q.data_ptr() + 4 * R, G * R, None, 0, 0, None, None, None, None, self._gamma,
                            lp[2].data_ptr(), lp[3].data_ptr(), N, A, la.buf.data_ptr(), self._conservative_weight,
                            self._alpha_threshold, None, 0, self.sums_ptr(S_ALPHA), done[0].data_ptr(), B, E, inv_b, 1,
                            self.counter_ptr(C_ALPHA), self._alpha_learning_rate, self.metric_ptr(M_ALPHA_LOSS),
                            self.metric_ptr(M_ALPHA), st)
        elif do_alpha:
            L.critic_loss(q.data_ptr() + 4 * R, G * R, None, 0, 0, None, None, None, None, self._gamma, lp[2].data_ptr(),
                          lp[3].data_ptr(), N, A, la.ptr("p"), self._conservative_weight, None, 0,
                          self.sums_ptr(S_ALPHA), None, B, E, inv_b, 0, st)
            if px is not None:
                # exchange + temperature step + alpha step in one launch
                L.dp_scalar_steps(self.sums_ptr(S_ALPHA), px.xchg_ptrs, px.flags_ptrs, px.world, px.rank, 1,
                                  self.counter_ptr(C_DRAW), lt.buf.data_ptr() if temp_merged else None,
                                  self.counter_ptr(C_TEMP), self._temp_learning_rate, self.metric_ptr(M_TEMP_LOSS),
                                  self.metric_ptr(M_TEMP), la.buf.data_ptr(), self.counter_ptr(C_ALPHA),
                                  self._alpha_learning_rate, inv_b / E, self._conservative_weight,
                                  self._alpha_threshold, self.metric_ptr(M_ALPHA_LOSS), self.metric_ptr(M_ALPHA), st)
            else:
                self._small_allreduce(px, self._slots[32 + S_ALPHA:32 + S_ALPHA + (4 if temp_merged else 3)], 1)
                if temp_merged:
                    temp_adam()
                L.cql_finalize(self.sums_ptr(S_ALPHA), la.ptr("p"), inv_b, E, self._conservative_weight,
                               self._alpha_threshold, 1, 1, self.metric_ptr(M_ALPHA_LOSS), la.ptr("g"), st)
                L.scalar_adam(la.ptr("p"), la.ptr("g"), la.ptr("m"), la.ptr("v"), self.counter_ptr(C_ALPHA),
                              self._alpha_learning_rate, 0.9, 0.999, 1e-8, self.metric_ptr(M_ALPHA), st)
        L.stream_join(st, side)
        dq = self.ws("is2_dq", E, R)
        if not dp:
            L.cql_loss_step(q.data_ptr(), GR, None if soft else q_t.data_ptr(), B, E,
                            q_tpn.data_ptr() if soft else None, db.ptr("rew"), db.ptr("term"), db.ptr("nsteps"),
                            self._gamma, lp[0].data_ptr(), lp[1].data_ptr(), N, A, la.buf.data_ptr(),
                            self._conservative_weight, self._alpha_threshold, dq.data_ptr(), R,
                            self.sums_ptr(S_CRITIC), done[1].data_ptr(), B, E, inv_b, 0, None, 0.0,
                            self.metric_ptr(M_CRITIC), None, st)
        else:
            L.critic_loss(q.data_ptr(), GR, None if soft else q_t.data_ptr(), B, E,
                          q_tpn.data_ptr() if soft else None, db.ptr("rew"), db.ptr("term"), db.ptr("nsteps"),
                          self._gamma, lp[0].data_ptr(), lp[1].data_ptr(), N, A, la.ptr("p"), self._conservative_weight,
                          dq.data_ptr(), R, self.sums_ptr(S_CRITIC), None, B, E, inv_b, 1, st)
        q_backward(0, R, ctx, dq, st)
        # the loss partial sums only feed the reported metric: with the peer exchange they ride along with the
        # gradient all-reduce inside the Adam kernel (no rendezvous of their own)
        csum = self._slots[32 + S_CRITIC:32 + S_CRITIC + 3]
        if dp and px is None:
            self._allreduce(csum)
            self._allreduce(q_net.arena.grads)
        q_net.adam(self._critic_learning_rate, st, tau=self._tau, peer=self._peer_args(px, q_net, (csum, 2)))
        # actor step on the updated critics
        ctx_a = q_net.ctx("aq", B, E, True)
        qa = self.ws("aq_q", E, B)
        q_forward("params", a_row0, B, ctx_a, qa, st)
        dqa = self.ws("a_dq", E, B)
        if not dp:
            L.sac_actor_step(qa.data_ptr(), B, lpm[1].data_ptr(), lt.ptr("p"), dqa.data_ptr(), B,
                             self.sums_ptr(S_ACTOR), done[2].data_ptr(), self.metric_ptr(M_ACTOR), B, E, inv_b, st)
        else:
            L.sac_actor_loss(qa.data_ptr(), B, lpm[1].data_ptr(), lt.ptr("p"), dqa.data_ptr(), B,
                             self.sums_ptr(S_ACTOR), B, E, inv_b, st)
        dxa = self.ws("a_dx", E, B, A)
        q_backward(a_row0, B, ctx_a, dqa, st, weight_grads=False, dx=dxa, lddx=A, stride_dx=B * A, dx_col0=O,
                   dx_cols=A)
        dhead = self.ws("pi_dhead", 1, B, 2 * A)
        L.sac_actor_backward(head.data_ptr(), 2 * A, nv("actor"), dxa.data_ptr(), A, B * A, E, lt.ptr("p"),
                             dhead.data_ptr(), 2 * A, B, A, MIN_LOGSTD, MAX_LOGSTD, inv_b, st)
        self._policy_backward_rows(db, acts_p, dhead, B)
        asum = self._slots[32 + S_ACTOR:32 + S_ACTOR + 1]
        if dp and px is None:
            self._allreduce(asum)
            self._allreduce(self._policy.arena.grads)
        self._policy.adam(self._actor_learning_rate, st, tau=self._tau,
                          peer=self._peer_args(px, self._policy, (asum, 3)))
        if dp:
            # metrics from the all-reduced sums, after everything that is on the critical path
            L.cql_finalize(self.sums_ptr(S_CRITIC), la.ptr("p"), inv_b, E, self._conservative_weight,
                           self._alpha_threshold, 0, 1 if N > 0 else 0, self.metric_ptr(M_CRITIC), None, st)
            # the actor metric is the all-reduced sum itself: update_fused_async reads it from the sums slot

    # ---- NVLink peer-memory exchange (csrc/comm.cu): all-reduce fused into the Adam pass, no NCCL in the update
    def _peer_setup(self):
        from ... import parallel

        px = parallel.new_peers(self._device)
        if px is not None and getattr(self._q_func, "_peer", None) is None:
            for net in (self._q_func, self._policy):   # collective: same order on every rank
                net._peer = px.register_arena(net.arena.grads)
        return px

    def _peer_args(self, px, net, small=None):
        if px is None:
            return None
        gptrs, fidx, cptr, gred, cptr2 = net._peer
        return (px, gptrs, fidx, cptr, self.counter_ptr(C_DRAW), small, gred, cptr2)

    def _small_allreduce(self, px, t, channel: int):
        if px is None:
            self._allreduce(t)
        else:
            self._lib.peer_allreduce_small(t.data_ptr(), t.numel(), px.xchg_ptrs, px.flags_ptrs, px.world, px.rank,
                                           channel, self.counter_ptr(C_DRAW), self._stream)

    # ------------------------------------------------------------------ fused update (CQL._update, cql.py:234-258)
    def update_fused(self, batch):
        names = self.update_fused_async(batch)
        return self._metrics_dict(names)

    def update_fused_async(self, batch):
        """Enqueue one whole update (no host sync); returns the metric slot names."""
        db = self.load_batch(batch, defer=True)
        do_temp, do_alpha = self._temp_learning_rate > 0, self._alpha_learning_rate > 0

        def program():
            ticks = [C_DRAW, C_CRITIC, C_ACTOR] + ([C_TEMP] if do_temp else []) + ([C_ALPHA] if do_alpha else [])
            self._tick(*ticks)
            self.zero_slots()
            self.fill_noise(db.B)
            acts_p, head = self._p_policy(db)
            if do_temp:
                self._p_temp(db, head)
            if do_alpha:
                self._p_alpha(db, head)
            q_t, q_tpn = self._p_target(db, head)
            self._p_critic(db, head, q_t=q_t, q_tpn=q_tpn, conservative=self._n_action_samples > 0)
            self._p_actor(db, acts_p, head)

        # fused glue (csrc/cql_fused.cu): bf16 mode over the whole-network kernels (any world size), fp32 mode over the
        # per-layer GEMMs on one GPU (the NCCL data-parallel fp32 path keeps the generic program)
        fused = self.fused_glue and ((self._precision == "bf16" and self._q_func.fused_ok and self._policy.fused_ok)
                                     or (self._precision == "fp32" and self.world_size == 1
                                         and not self._q_func.wide_head))
        if fused and self.world_size > 1 and not hasattr(self, "_px"):
            self._px = self._peer_setup()  # collective IPC rendezvous: outside the dry pass / graph capture
        self.run_program((type(self).__name__, db.B, do_temp, do_alpha, self._noise_injected, fused),
                         (lambda: self._program_fused(db, do_temp, do_alpha)) if fused else program)
        names = []
        if do_temp:
            names += [(M_TEMP_LOSS, "temp_loss"), (M_TEMP, "temp")]
        if do_alpha:
            names += [(M_ALPHA_LOSS, "alpha_loss"), (M_ALPHA, "alpha")]
        names += [(M_CRITIC, "critic_loss"), ((32 + S_ACTOR) if fused and self.world_size > 1 else M_ACTOR, "actor_loss")]
        return names

    # ------------------------------------------------------------------ reference hooks (eager, one sync each)
    def _begin(self, batch, *ticks):
        db = self.load_batch(batch)
        if ticks:
            self._tick(*ticks)
        self.zero_slots()
        self.fill_noise(db.B)
        return db

    def update_temp(self, batch):
        db = self._begin(batch, C_DRAW, C_TEMP)
        _, head = self._p_policy(db)
        self._p_temp(db, head)
        v = self.read_slots()
        return v[M_TEMP_LOSS].copy(), v[M_TEMP].copy()

    def update_alpha(self, batch):
        db = self._begin(batch, C_DRAW, C_ALPHA)
        _, head = self._p_policy(db)
        self._p_alpha(db, head)
        v = self.read_slots()
        return v[M_ALPHA_LOSS].copy(), v[M_ALPHA].copy()

    def compute_target(self, batch) -> torch.Tensor:
        db = self._begin(batch)
        _, head = self._p_policy(db)
        q_t, q_tpn = self._p_target(db, head)
        self.sync()
        return (q_t.min(dim=0).values if q_tpn is None else q_tpn).view(-1, 1).clone()

    def compute_critic_loss(self, batch, q_tpn: torch.Tensor) -> torch.Tensor:
        db = self._begin(batch)
        _, head = self._p_policy(db)
        self._p_critic(db, head, q_tpn=q_tpn.to(self._device).reshape(-1).contiguous(), backward=False)
        self.sync()
        return self._slots[M_CRITIC].clone()

    def _compute_conservative_loss(self, obs_t=None, act_t=None, obs_tp1=None, batch=None) -> torch.Tensor:
        """cql_impl.py:196-223.  Accepts either the reference's three tensors or a minibatch."""
        if batch is None:
            batch = _TensorBatch(obs_t, act_t, obs_tp1)
        db = self._begin(batch)
        _, head = self._p_policy(db)
        self._p_critic(db, head, q_tpn=self.ws("zero_tpn", db.B), backward=False, td=False)
        self.sync()
        return self._slots[M_CRITIC].clone()

    def update_critic(self, batch) -> np.ndarray:
        db = self._begin(batch, C_DRAW, C_CRITIC)
        _, head = self._p_policy(db)
        q_t, q_tpn = self._p_target(db, head)
        self._p_critic(db, head, q_t=q_t, q_tpn=q_tpn, sync_target=False)
        return self.read_slots()[M_CRITIC].copy()

    def compute_actor_loss(self, batch) -> torch.Tensor:
        """SACImpl.compute_actor_loss (sac_impl.py:114-121): mean(exp(log_temp) * log pi(a|s) - min_e Q_e(s, a)) with
        a freshly sampled a ~ pi(.|s); nothing is stepped."""
        db = self._begin(batch)
        acts_p, head = self._p_policy(db)
        self._p_actor(db, acts_p, head, step=False)
        self.sync()
        return self._slots[M_ACTOR].clone()

    def update_actor(self, batch) -> np.ndarray:
        db = self._begin(batch, C_DRAW, C_ACTOR)
        acts_p, head = self._p_policy(db)
        self._p_actor(db, acts_p, head, sync_target=False)
        return self.read_slots()[M_ACTOR].copy()


    def _checkpoint_views(self):
        v = super()._checkpoint_views()
        v.update({"_log_temp": self._log_temp, "_log_alpha": self._log_alpha,
                  "_temp_optim": self._log_temp.optim_view(self._temp_learning_rate),
                  "_alpha_optim": self._log_alpha.optim_view(self._alpha_learning_rate)})
        return v

    # ---- evaluation API
    def _predict_best_action(self, obs: torch.Tensor) -> torch.Tensor:
        """SquashedNormalPolicy.best_action = tanh(mu) (policies.py:247-249)."""
        head = self._policy_head(obs, head_tanh=False)
        return torch.tanh(head[0, :, :self._action_size])

    def sample_action(self, x) -> np.ndarray:
        """policy.sample (policies.py:167-181): tanh(mu + exp(clamp(logstd)) eps), eps ~ N(0, 1)."""
        obs = self._eval_obs(x)
        n, A = obs.shape[0], self._action_size
        head = self._policy_head(obs, head_tanh=False)
        with torch.cuda.stream(self._stream_obj):
            eps = torch.randn(1, n, A, device=self._device)
        act = self.ws("eval_act", n, A)
        self._lib.policy_sample_rows(head.data_ptr(), 2 * A, eps.data_ptr(), None, 0, None, 0, act.data_ptr(), None, n,
                                     1, 0, A, MIN_LOGSTD, MAX_LOGSTD, 0, self._stream)
        self.unscale_actions(act)   # algos/torch/base.py:77-79
        self.sync()
        return act.detach().cpu().numpy()


class _TensorBatch:
    """Adapter so the three-tensor `_compute_conservative_loss(obs_t, act_t, obs_tp1)` signature can
    reuse the minibatch staging path."""

    def __init__(self, obs_t, act_t, obs_tp1):
        n = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
        self.observations, self.actions, self.next_observations = n(obs_t), n(act_t), n(obs_tp1)
        B = self.observations.shape[0]
        self.rewards = np.zeros((B, 1), np.float32)
        self.terminals = np.zeros((B, 1), np.float32)
        self.n_steps = np.ones((B, 1), np.float32)
