"""GPU: observation / action / reward scalers on the update path -- the scaling kernels
and whole updates with scalers configured, against tests/golden/scalers.npz written by the unmodified reference.
Tolerances: bit-exact where the kernel restates the reference's float32 operator sequence (min-max observation and
action scaling, action un-scaling), 1e-6 for reward scalers (python-double constants rounded once), 1e-5 (fp32 mode)
on update metrics and post-step parameters."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests.golden_io import Case, load_scalers
from tests.test_update_gpu import REL, _assert_metrics, _assert_params, _ns

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda:0")


def test_scaling_kernels_match_reference_transforms():
    from d3rlpy_b200 import preprocessing as pp
    from d3rlpy_b200._lib import lib

    L, z = lib(), load_scalers()
    st = torch.cuda.current_stream().cuda_stream
    x, a, r, u = (z[k] for k in ("tr/batch/observations", "tr/batch/actions", "tr/batch/rewards", "tr/unit_actions"))

    mm = pp.MinMaxScaler(minimum=z["fit/min_max/minimum"], maximum=z["fit/min_max/maximum"])
    sub, div, eps = mm.affine_f32()
    d, sub_d, div_d = _dev(x), _dev(sub), _dev(div)
    L.standardize(d.data_ptr(), sub_d.data_ptr(), div_d.data_ptr(), eps, x.shape[0], x.shape[1], st)
    assert np.array_equal(d.cpu().numpy(), z["tr/min_max"])

    am = pp.MinMaxActionScaler(minimum=z["fit/action_min_max/minimum"], maximum=z["fit/action_min_max/maximum"])
    mn, mx = (_dev(v) for v in am.bounds_f32())
    d = _dev(a)
    L.scale_actions(d.data_ptr(), mn.data_ptr(), mx.data_ptr(), a.shape[0], a.shape[1], st)
    assert np.array_equal(d.cpu().numpy(), z["tr/action_min_max"])
    d = _dev(u)
    L.unscale_actions(d.data_ptr(), mn.data_ptr(), mx.data_ptr(), u.shape[0], u.shape[1], st)
    assert np.array_equal(d.cpu().numpy(), z["tr/action_min_max_reverse"])

    p = lambda name, key: float(z[f"fit/reward_{name}/{key}"])
    scalers = {"multiply": pp.MultiplyRewardScaler(multiplier=0.25),
               "clip": pp.ClipRewardScaler(-1.0, 1.5, multiplier=2.0),
               "min_max": pp.MinMaxRewardScaler(minimum=p("min_max", "minimum"), maximum=p("min_max", "maximum"),
                                                multiplier=3.0),
               "standard": pp.StandardRewardScaler(mean=p("standard", "mean"), std=p("standard", "std"), multiplier=0.5),
               "return": pp.ReturnBasedRewardScaler(return_max=p("return", "return_max"),
                                                    return_min=p("return", "return_min"), multiplier=1000.0)}
    for name, s in scalers.items():
        d = _dev(r)
        L.scale_rewards(d.data_ptr(), d.numel(), *s.constants(), st)
        assert np.allclose(d.cpu().numpy(), z[f"tr/reward_{name}"], rtol=1e-6, atol=1e-7), name
    # NaN rewards stay NaN through the clip (torch.clamp semantics); empty input is a no-op
    d = torch.tensor([float("nan"), 5.0, -5.0], device="cuda:0")
    L.scale_rewards(d.data_ptr(), 3, -1.0, 1.0, 0.0, 2.0, 1.0, st)
    got = d.cpu().numpy()
    assert np.isnan(got[0]) and got[1] == 2.0 and got[2] == -2.0
    L.scale_rewards(None, 0, -1.0, 1.0, 0.0, 1.0, 1.0, st)


def _build(z, name, **kw):
    from d3rlpy_b200 import preprocessing as pp
    from d3rlpy_b200.algos import CQL, DoubleDQN, TD3PlusBC

    case = Case(z, name)
    c = case.cfg
    O, A, B = int(c["obs"]), int(c["act"]), int(c["batch"])
    if name == "td3bc_scaled":
        algo = TD3PlusBC(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=B, n_steps=2,
                         scaler=pp.MinMaxScaler(minimum=z[f"{name}/obs_minimum"], maximum=z[f"{name}/obs_maximum"]),
                         action_scaler=pp.MinMaxActionScaler(minimum=z[f"{name}/act_minimum"],
                                                             maximum=z[f"{name}/act_maximum"]),
                         reward_scaler=pp.StandardRewardScaler(mean=c["reward_mean"], std=c["reward_std"],
                                                               eps=c["reward_eps"], multiplier=c["reward_multiplier"]),
                         **kw)
        groups = ("q", "pi", "targ_q", "targ_pi")
    elif name == "cql_scaled":
        algo = CQL(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=B,
                   n_action_samples=int(c["n_action_samples"]),
                   scaler=pp.StandardScaler(mean=z[f"{name}/obs_mean"], std=z[f"{name}/obs_std"]),
                   action_scaler=pp.MinMaxActionScaler(minimum=z[f"{name}/act_minimum"],
                                                       maximum=z[f"{name}/act_maximum"]),
                   reward_scaler=pp.ClipRewardScaler(c["reward_low"], c["reward_high"], c["reward_multiplier"]), **kw)
        groups = ("q", "pi", "targ_q", "log_temp", "log_alpha")
    else:
        algo = DoubleDQN(encoder_factory=[32, 32], batch_size=B, target_update_interval=2,
                         scaler=pp.MinMaxScaler(minimum=z[f"{name}/obs_minimum"], maximum=z[f"{name}/obs_maximum"]),
                         reward_scaler=pp.ReturnBasedRewardScaler(return_max=c["return_max"],
                                                                  return_min=c["return_min"],
                                                                  multiplier=c["reward_multiplier"]), **kw)
        groups = ("q", "targ_q")
    algo.create_impl((O,), A)
    impl = algo.impl
    impl.q_function.load_state_dict(case.group("init", "q"))
    impl.targ_q_function.load_state_dict(case.group("init", "q"))
    if "pi" in groups:
        impl.policy.load_state_dict(case.group("init", "pi"))
        impl.targ_policy.load_state_dict(case.group("init", "pi"))
    return case, algo, groups


def _view(impl, grp):
    return {"q": lambda: impl.q_function, "pi": lambda: impl.policy, "targ_q": lambda: impl.targ_q_function,
            "targ_pi": lambda: impl.targ_policy, "log_temp": lambda: impl._log_temp,
            "log_alpha": lambda: impl._log_alpha}[grp]()


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("name", ["td3bc_scaled", "cql_scaled", "dqn_scaled"])
def test_update_with_scalers_matches_reference_golden(name, use_graph):
    z = load_scalers()
    case, algo, groups = _build(z, name)
    impl = algo.impl
    impl.use_graph = use_graph
    B = int(case.cfg["batch"])
    for s in range(case.steps):
        if case.noise(s):
            impl.inject_noise(case.noise(s), B)
        m = algo.update(_ns(case.batch(s)))
        _assert_metrics(m, case.step_metrics(s), f"{name} step {s}")
    for grp in groups:
        _assert_params(_view(impl, grp).state_dict(), case.group("final", grp), f"{name}/{grp}")
    # evaluation path: raw observations in, actions back in the data's own range (algos/torch/base.py:50-80)
    if name != "dqn_scaled":
        got = algo.predict(z[f"{name}/eval_x"])
        ref = z[f"{name}/predict"]
        assert np.abs(got - ref).max() <= REL * max(1.0, np.abs(ref).max()) * 10, name
    if name == "td3bc_scaled":
        got = algo.predict_value(z[f"{name}/eval_x"], z[f"{name}/eval_action"])
        ref = z[f"{name}/predict_value"]
        assert np.abs(got - ref).max() <= REL * max(1.0, np.abs(ref).max()) * 10


def test_bf16_update_with_scalers_within_tolerance():
    z = load_scalers()
    case, algo, groups = _build(z, "cql_scaled", precision="bf16")
    B = int(case.cfg["batch"])
    for s in range(case.steps):
        algo.impl.inject_noise(case.noise(s), B)
        m = algo.update(_ns(case.batch(s)))
        _assert_metrics(m, case.step_metrics(s), f"bf16 step {s}", rel=1e-2)


def test_device_gathered_batches_are_scaled_once():
    """`algo.update(TransitionMiniBatch(transitions))` (the reference's call, base.py:746-758) and `from_indices`
    batches must see exactly the transforms a host batch sees; the caller's minibatch keeps showing raw data."""
    from d3rlpy_b200.dataset import MDPDataset, TransitionMiniBatch

    z = load_scalers()
    ds = MDPDataset(z["data/observations"], z["data/actions"], z["data/rewards"], z["data/terminals"],
                    z["data/episode_terminals"])
    trs = ds.transitions()
    rs = np.random.RandomState(3)
    idx = [rs.randint(len(trs), size=16) for _ in range(3)]
    runs = {}
    for mode in ("host", "ctor", "from_indices"):
        case, algo, groups = _build(z, "td3bc_scaled")
        impl = algo.impl
        out = []
        for s, ix in enumerate(idx):
            impl.inject_noise(case.noise(s), 16)
            if mode == "from_indices":
                batch = TransitionMiniBatch.from_indices(ds.device_replay(impl._device), ix, n_steps=2, gamma=0.99,
                                                         scaler=algo.scaler)
                assert batch.scaled == {"obs"}
            else:
                batch = TransitionMiniBatch([trs[i] for i in ix], n_steps=2, gamma=0.99)
            raw = {k: np.array(getattr(batch, k)) for k in ("observations", "actions", "rewards")}
            if mode == "host":
                batch = SimpleNamespace(**{k: np.array(getattr(batch, k)) for k in (
                    "observations", "actions", "rewards", "next_observations", "terminals", "n_steps")})
            out.append(algo.update(batch))
            if mode == "ctor":   # the caller's buffers are untouched; repeating the update on them scales once again
                for k, v in raw.items():
                    assert np.array_equal(np.array(getattr(batch, k)), v), k
        runs[mode] = (out, {g: {k: v.clone() for k, v in _view(impl, g).state_dict().items()} for g in groups})
    for mode in ("ctor", "from_indices"):
        for a, b in zip(runs["host"][0], runs[mode][0]):
            _assert_metrics(b, {k: float(v) for k, v in a.items()}, mode, rel=1e-6)
        for g in runs["host"][1]:
            _assert_params(runs[mode][1][g], runs["host"][1][g], f"{mode}/{g}", rel=1e-6)


def test_fit_fits_and_applies_scalers():
    """fit(): scalers are fitted on the data (base.py:566-585) and applied inside the device loop; the same index
    stream fed through host batches gives the same metrics."""
    from d3rlpy_b200.algos import DDPG
    from d3rlpy_b200.algos.base import random_iterator_indices
    from d3rlpy_b200.dataset import MDPDataset, TransitionMiniBatch

    z = load_scalers()
    ds = MDPDataset(z["data/observations"], z["data/actions"], z["data/rewards"], z["data/terminals"],
                    z["data/episode_terminals"])
    kw = dict(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=16, scaler="min_max",
              action_scaler="min_max", reward_scaler="standard")
    a = DDPG(**kw)
    hist = a.fit(ds, n_steps=6, n_steps_per_epoch=3, seed=5)
    assert np.array_equal(a.scaler._minimum, z["fit/min_max/minimum"])
    assert np.array_equal(a.action_scaler._maximum, z["fit/action_min_max/maximum"])
    assert abs(a.reward_scaler._mean - float(z["fit/reward_standard/mean"])) < 1e-9
    b = DDPG(**kw)
    for sc in (b.scaler, b.action_scaler, b.reward_scaler):
        sc.fit(ds)
    b.build_with_dataset(ds)
    rng = np.random.RandomState(5)
    trs = ds.transitions()
    for epoch in range(2):
        acc = {}
        for ix in random_iterator_indices(rng, len(trs), 3, 16):
            mb = TransitionMiniBatch([trs[i] for i in ix])
            host = SimpleNamespace(**{k: np.array(getattr(mb, k)) for k in (
                "observations", "actions", "rewards", "next_observations", "terminals", "n_steps")})
            for k, v in b.update(host).items():
                acc.setdefault(k, []).append(float(v))
        for k, v in acc.items():
            assert abs(hist[epoch][1][k] - np.mean(v)) <= 1e-5 * max(1.0, abs(np.mean(v))), (epoch, k)
    # pixels + reward clipping (the Atari reproduction's pairing): rewards are scaled, frames left to the conv load
    from d3rlpy_b200.algos import DQN
    from d3rlpy_b200.preprocessing import ClipRewardScaler

    rs = np.random.RandomState(0)
    n = 64
    frames = rs.randint(0, 256, size=(n, 1, 84, 84)).astype(np.uint8)
    pix = MDPDataset(frames, rs.randint(0, 4, size=n).astype(np.int32), (rs.randn(n) * 5).astype(np.float32),
                     (np.arange(n) % 32 == 31).astype(np.float32), discrete_action=True)
    dq = DQN(batch_size=8, n_frames=4, scaler="pixel", reward_scaler=ClipRewardScaler(-1.0, 1.0))
    dq.build_with_dataset(pix)
    mb = TransitionMiniBatch(pix.transitions()[:8], n_frames=4)
    raw_rewards = np.array(mb.rewards)
    m = dq.update(mb)
    assert np.isfinite(m["loss"])
    seen = dq.impl._batch.view("rew").cpu().numpy()
    assert np.array_equal(seen, np.clip(raw_rewards, -1.0, 1.0))
    assert np.array_equal(np.array(mb.rewards), raw_rewards)


def test_fit_over_an_episode_list_trains_on_those_transitions_only():
    """`fit(train_episodes)` (base.py:494-507), the train/test-split pattern: the index stream runs over the list's
    transitions; the same stream walked by hand through host minibatches gives the same metrics."""
    from d3rlpy_b200.algos import DDPG
    from d3rlpy_b200.algos.base import random_iterator_indices
    from d3rlpy_b200.dataset import MDPDataset, TransitionMiniBatch

    z = load_scalers()
    ds = MDPDataset(z["data/observations"], z["data/actions"], z["data/rewards"], z["data/terminals"],
                    z["data/episode_terminals"])
    train = ds.episodes[1:5]
    kw = dict(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=16, scaler="standard",
              reward_scaler="min_max", n_steps=2)
    a = DDPG(**kw)
    hist = a.fit(train, n_steps=4, n_steps_per_epoch=2, seed=9)
    trs = [t for e in train for t in e.transitions]
    obs = np.stack([t.observation for t in trs]).astype(np.float64)
    assert np.allclose(a.scaler._mean, obs.mean(0), rtol=1e-12) and a.reward_scaler._minimum == min(t.reward for t in trs)
    b = DDPG(**kw)
    for sc in (b.scaler, b.reward_scaler):
        sc.fit(trs)
    b.build_with_dataset(ds)
    rng = np.random.RandomState(9)
    for epoch in range(2):
        acc = {}
        for ix in random_iterator_indices(rng, len(trs), 2, 16):
            mb = TransitionMiniBatch([trs[i] for i in ix], n_steps=2, gamma=0.99)
            host = SimpleNamespace(**{k: np.array(getattr(mb, k)) for k in (
                "observations", "actions", "rewards", "next_observations", "terminals", "n_steps")})
            for k, v in b.update(host).items():
                acc.setdefault(k, []).append(float(v))
        for k, v in acc.items():
            assert abs(hist[epoch][1][k] - np.mean(v)) <= 1e-5 * max(1.0, abs(np.mean(v))), (epoch, k)
