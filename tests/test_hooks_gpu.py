"""GPU: the torch impl hooks BASELINE.json's north_star names -- update_temp / update_alpha / update_critic /
update_actor / update_critic_target / update_actor_target, compute_target, compute_critic_loss, compute_actor_loss --
called one by one the way `X._update` sequences them (cql.py:234-258, td3_plus_bc.py:177-192, bcq.py:261-279) against
one oracle update on identical weights, minibatch and injected noise.  fp32 mode, tolerance 2e-5 on every returned
loss and on the post-step parameters; the `compute_*` hooks must return exactly what the matching `update_*` hook
reports (same kernels, nothing stepped in between)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import update as ou
from tests.test_update_gpu import _assert_params, _synthetic_batch

pytestmark = pytest.mark.gpu
REL = 2e-5


def _close(got, ref, what):
    got, ref = float(got), float(ref)
    assert abs(got - ref) <= REL * max(1.0, abs(ref)) + 1e-6, (what, got, ref)


def _snapshot(impl, views):
    return {name: {k: v.clone() for k, v in getattr(impl, name).state_dict().items()} for name in views}


def _unchanged(impl, before, what):
    for name, sd in before.items():
        for k, v in getattr(impl, name).state_dict().items():
            assert torch.equal(v, sd[k]), f"{what} modified {name}/{k}"


def test_cql_hooks_one_by_one_match_oracle_update():
    from d3rlpy_b200.algos import CQL

    O, A, B, N, H = 9, 4, 32, 5, [64, 64, 64]
    orc = ou.CQL(O, A, hidden=H, n_action_samples=N, seed=11)
    algo = CQL(actor_encoder_factory=H, critic_encoder_factory=H, batch_size=B, n_action_samples=N)
    algo.create_impl((O,), A)
    impl = algo.impl
    for view, p in ((impl.q_function, orc.q), (impl.targ_q_function, orc.q), (impl.policy, orc.pi),
                    (impl.targ_policy, orc.pi)):
        view.load_state_dict(p)
    arrays = _synthetic_batch(np.random.RandomState(2), B, O, A)
    noise = ou.Noise(seed=21)
    ref = orc.update(ou.Batch(arrays), noise)
    impl.inject_noise(noise.log, B)
    b = SimpleNamespace(**arrays)

    loss, temp = impl.update_temp(b)
    _close(loss, ref["temp_loss"], "temp_loss"), _close(temp, ref["temp"], "temp")
    loss, alpha = impl.update_alpha(b)
    _close(loss, ref["alpha_loss"], "alpha_loss"), _close(alpha, ref["alpha"], "alpha")

    before = _snapshot(impl, ("q_function", "policy", "targ_q_function", "targ_policy"))
    q_tpn = impl.compute_target(b)
    assert tuple(q_tpn.shape) == (B, 1)
    dry = impl.compute_critic_loss(b, q_tpn)
    _unchanged(impl, before, "compute_target / compute_critic_loss")
    loss = impl.update_critic(b)
    _close(loss, ref["critic_loss"], "critic_loss")
    _close(dry, loss, "compute_critic_loss vs update_critic")

    before = _snapshot(impl, ("q_function", "policy"))
    dry = impl.compute_actor_loss(b)
    _unchanged(impl, before, "compute_actor_loss")
    loss = impl.update_actor(b)
    _close(loss, ref["actor_loss"], "actor_loss")
    _close(dry, loss, "compute_actor_loss vs update_actor")

    impl.update_critic_target()
    impl.update_actor_target()
    impl.sync()
    for grp, view, refp in (("q", impl.q_function, orc.q), ("pi", impl.policy, orc.pi),
                            ("targ_q", impl.targ_q_function, orc.targ_q), ("targ_pi", impl.targ_policy, orc.targ_pi),
                            ("log_temp", impl._log_temp, orc.log_temp), ("log_alpha", impl._log_alpha, orc.log_alpha)):
        _assert_params(view.state_dict(), refp, grp, rel=REL)


@pytest.mark.parametrize("cloning", [True, False])
def test_td3_family_hooks_one_by_one_match_oracle_update(cloning):
    from d3rlpy_b200.algos import TD3, TD3PlusBC

    O, A, B, H = 7, 3, 32, [64, 64]
    orc = (ou.TD3PlusBC if cloning else ou.TD3)(O, A, hidden=H, seed=12)
    algo = (TD3PlusBC if cloning else TD3)(actor_encoder_factory=H, critic_encoder_factory=H, batch_size=B, scaler=None)
    algo.create_impl((O,), A)
    impl = algo.impl
    for view, p in ((impl.q_function, orc.q), (impl.targ_q_function, orc.q), (impl.policy, orc.pi),
                    (impl.targ_policy, orc.pi)):
        view.load_state_dict(p)
    arrays = _synthetic_batch(np.random.RandomState(3), B, O, A)
    noise = ou.Noise(seed=22)
    ref = orc.update(ou.Batch(arrays), noise)   # grad_step 0: critic, actor and both soft syncs
    impl.inject_noise(noise.log, B)
    b = SimpleNamespace(**arrays)

    before = _snapshot(impl, ("q_function", "policy", "targ_q_function", "targ_policy"))
    q_tpn = impl.compute_target(b)
    dry = impl.compute_critic_loss(b, q_tpn)
    _unchanged(impl, before, "compute_target / compute_critic_loss")
    loss = impl.update_critic(b)
    _close(loss, ref["critic_loss"], "critic_loss")
    _close(dry, loss, "compute_critic_loss vs update_critic")

    before = _snapshot(impl, ("q_function", "policy"))
    dry = impl.compute_actor_loss(b)
    _unchanged(impl, before, "compute_actor_loss")
    loss = impl.update_actor(b)
    _close(loss, ref["actor_loss"], "actor_loss")
    _close(dry, loss, "compute_actor_loss vs update_actor")

    impl.update_critic_target()
    impl.update_actor_target()
    impl.sync()
    for grp, view, refp in (("q", impl.q_function, orc.q), ("pi", impl.policy, orc.pi),
                            ("targ_q", impl.targ_q_function, orc.targ_q), ("targ_pi", impl.targ_policy, orc.targ_pi)):
        _assert_params(view.state_dict(), refp, grp, rel=REL)


def test_bcq_hooks_one_by_one_match_oracle_update():
    from d3rlpy_b200.algos import BCQ

    O, A, B, N, H, V = 8, 3, 32, 10, [64, 48], [96, 96]
    orc = ou.BCQ(O, A, hidden=H, vae_hidden=V, n_action_samples=N, seed=13)
    algo = BCQ(actor_encoder_factory=H, critic_encoder_factory=H, imitator_encoder_factory=V, batch_size=B,
               n_action_samples=N)
    algo.create_impl((O,), A)
    impl = algo.impl
    for view, p in ((impl.q_function, orc.q), (impl.targ_q_function, orc.q), (impl.policy, orc.pi),
                    (impl.targ_policy, orc.pi), (impl.imitator, orc.imitator)):
        view.load_state_dict(p)
    arrays = _synthetic_batch(np.random.RandomState(4), B, O, A)
    noise = ou.Noise(seed=23)
    ref = orc.update(ou.Batch(arrays), noise)
    impl.inject_noise(noise.log, B)
    b = SimpleNamespace(**arrays)

    _close(impl.update_imitator(b), ref["imitator_loss"], "imitator_loss")

    before = _snapshot(impl, ("q_function", "policy", "targ_q_function", "targ_policy"))
    q_tpn = impl.compute_target(b)
    assert tuple(q_tpn.shape) == (B, 1)
    dry = impl.compute_critic_loss(b, q_tpn)
    _unchanged(impl, before, "compute_target / compute_critic_loss")
    loss = impl.update_critic(b)
    _close(loss, ref["critic_loss"], "critic_loss")
    _close(dry, loss, "compute_critic_loss vs update_critic")

    before = _snapshot(impl, ("q_function", "policy"))
    dry = impl.compute_actor_loss(b)
    _unchanged(impl, before, "compute_actor_loss")
    loss = impl.update_actor(b)
    _close(loss, ref["actor_loss"], "actor_loss")
    _close(dry, loss, "compute_actor_loss vs update_actor")

    impl.update_actor_target()
    impl.update_critic_target()
    impl.sync()
    for grp, view, refp in (("q", impl.q_function, orc.q), ("pi", impl.policy, orc.pi),
                            ("imitator", impl.imitator, orc.imitator), ("targ_q", impl.targ_q_function, orc.targ_q),
                            ("targ_pi", impl.targ_policy, orc.targ_pi)):
        _assert_params(view.state_dict(), refp, grp, rel=REL)
