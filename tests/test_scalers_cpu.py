"""CPU: the scaler restatement (oracle/scalers.py) and the host side of the product scalers
(d3rlpy_b200/preprocessing.py: fit, constants handed to the kernels, params.json encoding) against
tests/golden/scalers.npz, which the unmodified reference wrote (tests/golden/make_golden_scalers.py)."""
import json

import numpy as np
import pytest
import torch

from oracle import scalers as osc
from oracle import update as ou
from tests.golden_io import Case, load_scalers


def _transition_view(z):
    """Transition arrays of the fixture's dataset, restated from _to_transitions (dataset.pyx:70-116): a timed-out
    episode drops its last step."""
    t, ept = z["data/terminals"], z["data/episode_terminals"]
    keep, ep, e0 = [], [], 0
    for i, e in enumerate(np.nonzero(ept)[0]):
        last = e if t[e] else e - 1
        keep += list(range(e0, last + 1))
        ep += [i] * (last + 1 - e0)
        e0 = e + 1
    keep = np.array(keep)
    return z["data/observations"][keep], z["data/actions"][keep], z["data/rewards"][keep], np.array(ep)


REWARD = {
    "multiply": (lambda: osc.MultiplyRewardScaler(0.25), dict(multiplier=0.25)),
    "clip": (lambda: osc.ClipRewardScaler(-1.0, 1.5, 2.0), dict(low=-1.0, high=1.5, multiplier=2.0)),
    "min_max": (lambda: osc.MinMaxRewardScaler(multiplier=3.0), dict(multiplier=3.0)),
    "standard": (lambda: osc.StandardRewardScaler(multiplier=0.5), dict(multiplier=0.5)),
    "return": (lambda: osc.ReturnBasedRewardScaler(multiplier=1000.0), dict(multiplier=1000.0)),
}


def test_oracle_scalers_match_reference_fit_and_transform():
    z = load_scalers()
    obs, act, rew, ep = _transition_view(z)
    x, a, r = (torch.tensor(z[f"tr/batch/{k}"]) for k in ("observations", "actions", "rewards"))
    mm = osc.MinMaxScaler().fit(obs)
    assert np.array_equal(mm.minimum, z["fit/min_max/minimum"]) and np.array_equal(mm.maximum, z["fit/min_max/maximum"])
    assert np.array_equal(mm(x).numpy(), z["tr/min_max"])                       # bit-exact: same float32 operators
    st = osc.StandardScaler().fit(obs)
    assert np.allclose(st.mean, z["fit/standard/mean"], rtol=1e-12, atol=0)
    assert np.allclose(st.std, z["fit/standard/std"], rtol=1e-12, atol=0)
    assert np.array_equal(st(x).numpy(), z["tr/standard"])
    am = osc.MinMaxActionScaler().fit(act)
    assert np.array_equal(am.minimum, z["fit/action_min_max/minimum"])
    assert np.array_equal(am(a).numpy(), z["tr/action_min_max"])
    assert np.array_equal(am.reverse(torch.tensor(z["tr/unit_actions"])).numpy(), z["tr/action_min_max_reverse"])
    rets = osc.episode_returns(rew, ep)
    for name, (make, _) in REWARD.items():
        s = make().fit(rew, rets)
        for k in [f for f in z.files if f.startswith(f"fit/reward_{name}/")]:
            assert abs(getattr(s, k.rsplit("/", 1)[1]) - float(z[k])) <= 1e-12 * max(1.0, abs(float(z[k]))), k
        assert np.allclose(s(r).numpy(), z[f"tr/reward_{name}"], rtol=1e-6, atol=1e-7), name


def test_product_scalers_fit_constants_and_json():
    """fit() over our MDPDataset gives the reference's parameters; the constants handed to the kernels, pushed through
    the kernels' formula in numpy float32, reproduce the reference's transforms; params.json round trip."""
    from d3rlpy_b200 import preprocessing as pp
    from d3rlpy_b200.algos.base import _scaler_from_json, _scaler_to_json
    from d3rlpy_b200.dataset import MDPDataset

    z = load_scalers()
    ds = MDPDataset(z["data/observations"], z["data/actions"], z["data/rewards"], z["data/terminals"],
                    z["data/episode_terminals"])
    x, a, r = (z[f"tr/batch/{k}"] for k in ("observations", "actions", "rewards"))
    f32 = np.float32

    mm = pp.MinMaxScaler(ds)
    assert np.array_equal(mm._minimum, z["fit/min_max/minimum"]) and np.array_equal(mm._maximum, z["fit/min_max/maximum"])
    sub, div, eps = mm.affine_f32()
    assert np.array_equal((x - sub) / (div + f32(eps)), z["tr/min_max"])
    st = pp.StandardScaler(ds)
    assert np.allclose(st._mean, z["fit/standard/mean"].reshape(-1), rtol=1e-12)
    assert np.allclose(st._std, z["fit/standard/std"].reshape(-1), rtol=1e-10)
    sub, div, eps = st.affine_f32()
    assert np.allclose((x - sub) / (div + f32(eps)), z["tr/standard"], rtol=1e-6, atol=1e-7)
    am = pp.MinMaxActionScaler(ds)
    assert np.array_equal(am._minimum, z["fit/action_min_max/minimum"])
    mn, mx = am.bounds_f32()
    assert np.array_equal(((a - mn) / (mx - mn)) * f32(2.0) - f32(1.0), z["tr/action_min_max"])
    u = z["tr/unit_actions"]
    assert np.array_equal(((mx - mn) * ((u + f32(1.0)) / f32(2.0))) + mn, z["tr/action_min_max_reverse"])

    made = {"multiply": pp.MultiplyRewardScaler(multiplier=0.25), "clip": pp.ClipRewardScaler(-1.0, 1.5, multiplier=2.0),
            "min_max": pp.MinMaxRewardScaler(ds, multiplier=3.0), "standard": pp.StandardRewardScaler(ds, multiplier=0.5),
            "return": pp.ReturnBasedRewardScaler(ds, multiplier=1000.0)}
    for name, s in made.items():
        assert s.get_type() == name and pp.REWARD_SCALER_LIST[name] is type(s)
        for k in [f for f in z.files if f.startswith(f"fit/reward_{name}/")]:
            got = s.get_params()[k.rsplit("/", 1)[1]]
            assert abs(got - float(z[k])) <= 1e-9 * max(1.0, abs(float(z[k]))), (k, got, float(z[k]))
        lo, hi, sb, mul, dv = (f32(c) for c in s.constants())
        got = (mul * (np.clip(r, lo, hi) - sb)) / dv            # reward_scale_kernel's formula (csrc/losses.cu)
        assert np.allclose(got, z[f"tr/reward_{name}"], rtol=1e-6, atol=1e-7), name
        doc = json.loads(json.dumps(_scaler_to_json(s)))
        back = _scaler_from_json(doc, pp.create_reward_scaler)
        assert type(back) is type(s) and back.constants() == s.constants()
    for s, create in ((mm, pp.create_scaler), (st, pp.create_scaler), (am, pp.create_action_scaler)):
        doc = json.loads(json.dumps(_scaler_to_json(s)))
        assert doc["type"] == s.TYPE
        back = _scaler_from_json(doc, create)
        assert type(back) is type(s)
        for k, v in s.get_params().items():
            assert np.allclose(np.asarray(back.get_params()[k], np.float64), np.asarray(v, np.float64), rtol=1e-15)
    # fitting from a list of Transitions (what the reference's fit() receives, base.py:566-585) == from the dataset
    trs = ds.transitions()
    sub = pp.ReturnBasedRewardScaler(multiplier=1000.0)
    sub.fit(trs[::7])   # whole-episode returns whichever transitions of an episode are passed
    assert sub.constants() == made["return"].constants()
    with pytest.raises(ValueError):
        _scaler_from_json({"type": "nope", "params": {}})


def test_algo_accepts_scaler_names_and_instances():
    from d3rlpy_b200 import preprocessing as pp
    from d3rlpy_b200.algos import CQL, DoubleDQN

    algo = CQL(scaler="min_max", action_scaler="min_max", reward_scaler="standard", use_gpu=None)
    assert isinstance(algo.scaler, pp.MinMaxScaler) and isinstance(algo.action_scaler, pp.MinMaxActionScaler)
    assert isinstance(algo.reward_scaler, pp.StandardRewardScaler)
    algo = DoubleDQN(scaler="pixel", reward_scaler=pp.ClipRewardScaler(-1.0, 1.0), use_gpu=None)
    assert isinstance(algo.scaler, pp.PixelScaler) and algo.reward_scaler.constants()[:2] == (-1.0, 1.0)
    with pytest.raises((ValueError, AssertionError)):
        CQL(reward_scaler="nope", use_gpu=None)


def _oracle_case(z, name):
    case = Case(z, name)
    c = case.cfg
    O, A = int(c["obs"]), int(c["act"])
    if name == "td3bc_scaled":
        algo = ou.TD3PlusBC(O, A, critics=case.group("init", "q"), policy=case.group("init", "pi"))
        sc = osc.MinMaxScaler(z[f"{name}/obs_minimum"], z[f"{name}/obs_maximum"])
        asc = osc.MinMaxActionScaler(z[f"{name}/act_minimum"], z[f"{name}/act_maximum"])
        rsc = osc.StandardRewardScaler(c["reward_mean"], c["reward_std"], c["reward_eps"], c["reward_multiplier"])
        groups = {"q": "q", "pi": "pi", "targ_q": "targ_q", "targ_pi": "targ_pi"}
    elif name == "cql_scaled":
        algo = ou.CQL(O, A, critics=case.group("init", "q"), policy=case.group("init", "pi"),
                      n_action_samples=int(c["n_action_samples"]))
        sc = osc.StandardScaler(z[f"{name}/obs_mean"], z[f"{name}/obs_std"])
        asc = osc.MinMaxActionScaler(z[f"{name}/act_minimum"], z[f"{name}/act_maximum"])
        rsc = osc.ClipRewardScaler(c["reward_low"], c["reward_high"], c["reward_multiplier"])
        groups = {"q": "q", "pi": "pi", "targ_q": "targ_q", "log_temp": "log_temp", "log_alpha": "log_alpha"}
    else:
        algo = ou.DiscreteCQL((O,), A, critics=case.group("init", "q"), double=True, conservative=False,
                              target_update_interval=2)
        sc = osc.MinMaxScaler(z[f"{name}/obs_minimum"], z[f"{name}/obs_maximum"])
        asc = None
        rsc = osc.ReturnBasedRewardScaler(c["return_max"], c["return_min"], c["reward_multiplier"])
        groups = {"q": "q", "targ_q": "targ_q"}
    return case, algo, sc, asc, rsc, groups


@pytest.mark.parametrize("name", ["td3bc_scaled", "cql_scaled", "dqn_scaled"])
def test_oracle_update_with_scalers_matches_reference(name):
    z = load_scalers()
    case, algo, sc, asc, rsc, groups = _oracle_case(z, name)
    for s in range(case.steps):
        m = algo.update(ou.Batch(case.batch(s), sc, rsc, asc), ou.Noise(injected=case.noise(s)))
        for k, v in case.step_metrics(s).items():
            assert abs(m[k] - v) <= 1e-5 * max(1.0, abs(v)), (name, s, k, m[k], v)
    for grp, attr in groups.items():
        for k, v in case.group("final", grp).items():
            got = getattr(algo, attr)[k].detach()
            assert float((got - v).abs().max()) <= 2e-6 * max(1.0, float(v.abs().max())), (name, grp, k)


def test_fit_accepts_episode_and_transition_lists():
    """LearnableBase.fit takes an MDPDataset, List[Episode] or List[Transition] (base.py:494-507); lists train on
    exactly their transitions.  Guards that fire before any device work are checked here, the training itself in
    tests/test_update_gpu.py."""
    from d3rlpy_b200 import preprocessing as pp
    from d3rlpy_b200.algos import CQL, DQN
    from d3rlpy_b200.dataset import MDPDataset

    z = load_scalers()
    ds = MDPDataset(z["data/observations"], z["data/actions"], z["data/rewards"], z["data/terminals"],
                    z["data/episode_terminals"])
    eps = ds.episodes
    sub = pp.TransitionSubset(eps[2:5])
    want = np.concatenate([[t._t for t in e.transitions] for e in eps[2:5]])
    assert sub._ds is ds and np.array_equal(sub._t_index, want) and len(sub) == len(want)
    trs = [t for e in eps[2:5] for t in e.transitions][::3]
    sub2 = pp.TransitionSubset(trs)
    assert np.array_equal(sub2._t_index, want[::3])
    # scalers fitted on a subset see only that subset (base.py:566-585 passes the transition list)
    mm = pp.MinMaxScaler()
    mm.fit(sub)
    obs = np.stack([t.observation for e in eps[2:5] for t in e.transitions])
    assert np.array_equal(mm._minimum.reshape(-1), obs.min(0)) and np.array_equal(mm._maximum.reshape(-1), obs.max(0))
    rb = pp.ReturnBasedRewardScaler()
    rb.fit(sub2)
    rets = [sum(t.reward for t in e.transitions) for e in eps[2:5]]
    assert abs(rb._return_max - max(rets)) < 1e-4 and abs(rb._return_min - min(rets)) < 1e-4
    with pytest.raises(ValueError, match="empty dataset"):
        CQL(use_gpu=None).fit([], n_steps=10, n_steps_per_epoch=10)
    with pytest.raises(ValueError, match="invalid dataset type"):
        CQL(use_gpu=None).fit([1, 2, 3], n_steps=10, n_steps_per_epoch=10)
    with pytest.raises(AssertionError, match="not compatible"):
        DQN(use_gpu=None).fit(eps[:2], n_steps=10, n_steps_per_epoch=10)
    with pytest.raises(ValueError, match="n_epochs or n_steps"):
        CQL(use_gpu=None).fit(ds)


def test_online_setup_fits_scalers_from_the_environment():
    """_setup_algo (online/iterators.py:76-96): min-max bounds come from the spaces; the standard scaler refuses."""
    from types import SimpleNamespace

    from d3rlpy_b200 import preprocessing as pp
    from d3rlpy_b200.online.iterators import _setup_algo

    env = SimpleNamespace(observation_space=SimpleNamespace(shape=(3,), low=np.array([-1, -2, -3], np.float32),
                                                            high=np.array([1, 2, 3], np.float32)),
                          action_space=SimpleNamespace(shape=(2,), low=np.array([-2, 0], np.float32),
                                                       high=np.array([2, 1], np.float32)))
    algo = SimpleNamespace(scaler=pp.MinMaxScaler(), action_scaler=pp.MinMaxActionScaler(), impl=object())
    _setup_algo(algo, env)
    assert algo.scaler._minimum.shape == (1, 3) and np.array_equal(algo.scaler._maximum[0], [1, 2, 3])
    mn, mx = algo.action_scaler.bounds_f32()
    assert np.array_equal(mn, [-2, 0]) and np.array_equal(mx, [2, 1])
    algo = SimpleNamespace(scaler=pp.StandardScaler(), action_scaler=None, impl=object())
    with pytest.raises(NotImplementedError):
        _setup_algo(algo, env)
    _setup_algo(SimpleNamespace(scaler=pp.PixelScaler(), action_scaler=None, impl=object()), env)


def test_fitter_epochs_callbacks_scorers_and_index_streams():
    """LearnableBase.fitter / fit (base.py:349-687) control flow with the device work stubbed out: `(epoch, metrics)`
    pairs, scorer entries merged into the metrics, `callback(algo, epoch, total_step)` after every step, the
    RandomIterator / RoundIterator index streams (restricted to the list's transitions when a list is given)."""
    from types import SimpleNamespace

    from d3rlpy_b200.algos import CQL
    from d3rlpy_b200.algos.base import random_iterator_indices, round_iterator_indices
    from d3rlpy_b200.dataset import MDPDataset

    z = load_scalers()
    ds = MDPDataset(z["data/observations"], z["data/actions"], z["data/rewards"], z["data/terminals"],
                    z["data/episode_terminals"])
    n_tr = ds._meta.shape[0]

    class Replay:
        def __len__(self):
            return n_tr

    ds.device_replay = lambda device=None: Replay()

    class Probe(CQL):
        def build_with_dataset(self, dataset):
            self._impl = SimpleNamespace(_device="cpu")

        def _fit_epoch(self, replay, idx, after_step=None):
            self.batches.append(np.array(idx))
            for _ in range(idx.shape[0]):
                self._grad_step += 1
                if after_step:
                    after_step()
            return {"loss": float(idx.shape[0])}

    algo = Probe(batch_size=8, use_gpu=None)
    algo.batches, seen = [], []
    out = algo.fit(ds, n_steps=12, n_steps_per_epoch=4, seed=3, eval_episodes=ds.episodes[:2],
                   scorers={"n_eval": lambda a, eps: float(len(eps))},
                   callback=lambda a, e, t: seen.append((e, t, a.grad_step)),
                   save_metrics=False, experiment_name="ignored", verbose=False, show_progress=False)
    assert [e for e, _ in out] == [1, 2, 3] and all(m == {"loss": 4.0, "n_eval": 2.0} for _, m in out)
    assert seen == [(1 + (t - 1) // 4, t, t) for t in range(1, 13)]
    rng = np.random.RandomState(3)
    for got in algo.batches:
        assert np.array_equal(got, random_iterator_indices(rng, n_tr, 4, 8))
    # generator form: work happens epoch by epoch, as the caller pulls
    algo = Probe(batch_size=8, use_gpu=None)
    algo.batches = []
    gen = algo.fitter(ds.episodes[1:4], n_epochs=2, seed=4)
    assert algo.batches == []
    epoch, metrics = next(gen)
    assert epoch == 1 and len(algo.batches) == 1
    assert [e for e, _ in gen] == [2] and len(algo.batches) == 2
    sub = np.concatenate([[t._t for t in e.transitions] for e in ds.episodes[1:4]])
    rng = np.random.RandomState(4)
    for got in algo.batches:
        want = sub[round_iterator_indices(rng, len(sub), 8, True)]
        assert np.array_equal(got, want) and set(got.reshape(-1)) <= set(sub)
