"""CPU: the oracle restatement vs the golden vectors produced by the live reference
(tests/golden/make_golden.py).  This is what pins the oracle where the reference is absent."""
import numpy as np
import pytest
import torch

from oracle import sampler as osampler
from oracle import update as oupdate
from tests.golden_io import Case, load_qr, load_sampler, load_siblings, load_update


def test_sampler_matches_reference_golden():
    z = load_sampler()
    for name in [str(c) for c in z["cases"]]:
        d = lambda k: z[f"{name}/data/{k}"]
        replay = osampler.FlatReplay(d("observations"), d("actions"), d("rewards"), d("terminals"), d("episode_terminals"))
        n_frames, n_steps = [int(v) for v in z[f"{name}/cfg"]]
        out = osampler.gather(replay, z[f"{name}/indices"], n_frames, n_steps, 0.99)
        for k, v in out.items():
            ref = z[f"{name}/ref/{k}"]
            assert ref.dtype == v.dtype and ref.shape == v.shape, (name, k)
            if k == "rewards":
                np.testing.assert_allclose(v, ref, rtol=1e-6, atol=1e-7, err_msg=f"{name}/{k}")
            else:
                assert np.array_equal(v, ref), (name, k)


def _build(case: Case):
    n, c = case.name, case.cfg
    g = lambda grp: case.group("init", grp)
    if n == "td3bc":
        algo = oupdate.TD3PlusBC(int(c["obs"]), int(c["act"]), critics=g("q"), policy=g("pi"))
        scaler = oupdate.standard_scaler(case.z["td3bc/scaler_mean"], case.z["td3bc/scaler_std"])
        return algo, scaler
    if n.startswith("cql"):
        return oupdate.CQL(int(c["obs"]), int(c["act"]), critics=g("q"), policy=g("pi"), n_action_samples=int(c["n"]),
                           soft_q_backup=bool(c["soft_q_backup"])), None
    if n == "bcq":
        return oupdate.BCQ(int(c["obs"]), int(c["act"]), critics=g("q"), policy=g("pi"), imitator=g("imitator"),
                           n_action_samples=int(c["n"])), None
    if n == "dcql_vec":
        return oupdate.DiscreteCQL((int(c["obs"]),), int(c["act"]), critics=g("q"),
                                   target_update_interval=int(c["interval"])), None
    if n == "dcql_pix":
        hw = int(c["hw"])
        return oupdate.DiscreteCQL((int(c["n_frames"]), hw, hw), int(c["act"]), critics=g("q")), oupdate.pixel_scaler()
    if n in ("qr_dcql_vec", "qr_dqn_vec"):
        plain = n == "qr_dqn_vec"
        return oupdate.DiscreteCQL((int(c["obs"]),), int(c["act"]), critics=g("q"), n_quantiles=int(c["n_quantiles"]),
                                   target_update_interval=int(c["interval"]), double=not plain,
                                   conservative=not plain), None
    if n in ("dqn_vec", "ddqn_vec", "nfq_vec"):
        return oupdate.DiscreteCQL((int(c["obs"]),), int(c["act"]), critics=g("q"), double=n == "ddqn_vec",
                                   target_update_interval=int(c["interval"]), conservative=False), None
    if n == "qr_dcql_pix":
        hw = int(c["hw"])
        return oupdate.DiscreteCQL((int(c["n_frames"]), hw, hw), int(c["act"]), critics=g("q"),
                                   n_quantiles=int(c["n_quantiles"])), oupdate.pixel_scaler()
    if n == "sac":
        return oupdate.SAC(int(c["obs"]), int(c["act"]), critics=g("q"), policy=g("pi")), None
    if n == "iql":
        return oupdate.IQL(int(c["obs"]), int(c["act"]), critics=g("q"), policy=g("pi"), value=g("v"),
                           max_weight=float(c["max_weight"])), None
    if n == "td3bc_qr":   # critics carry 8-quantile heads: the oracle switches on the head width
        return oupdate.TD3PlusBC(int(c["obs"]), int(c["act"]), critics=g("q"), policy=g("pi")), None
    if n == "ddpg_qr":
        return oupdate.DDPG(int(c["obs"]), int(c["act"]), critics=g("q"), policy=g("pi"), n_critics=2), None
    if n == "ddpg":
        return oupdate.DDPG(int(c["obs"]), int(c["act"]), critics=g("q"), policy=g("pi")), None
    if n == "td3":
        return oupdate.TD3(int(c["obs"]), int(c["act"]), critics=g("q"), policy=g("pi")), None
    raise KeyError(n)


@pytest.mark.parametrize("name", ["td3bc", "cql", "cql_softq", "bcq", "dcql_vec", "dcql_pix", "sac", "td3", "ddpg", "iql", "td3bc_qr", "ddpg_qr",
                                  "qr_dcql_vec", "qr_dqn_vec", "qr_dcql_pix", "dqn_vec", "ddqn_vec", "nfq_vec"])
def test_update_matches_reference_golden(name):
    torch.set_num_threads(1)
    z = load_siblings() if name in ("sac", "td3", "ddpg", "iql", "td3bc_qr", "ddpg_qr") else load_qr() if name.startswith("qr_") or name.endswith("dqn_vec") or name == "nfq_vec" else load_update()
    case = Case(z, name)
    algo, scaler = _build(case)
    for s in range(case.steps):
        m = algo.update(oupdate.Batch(case.batch(s), scaler), oupdate.Noise(injected=case.noise(s)))
        ref = case.step_metrics(s)
        assert set(m) == set(ref)
        for k in ref:
            assert abs(m[k] - ref[k]) <= 1e-5 * max(1.0, abs(ref[k])), (name, s, k, m[k], ref[k])
    groups = {"q": "q", "pi": "pi", "v": "v", "targ_q": "targ_q", "targ_pi": "targ_pi", "imitator": "imitator",
              "log_temp": "log_temp", "log_alpha": "log_alpha"}
    for grp, attr in groups.items():
        ref = case.group("final", grp)
        if not ref:
            continue
        got = getattr(algo, attr)
        for k, v in ref.items():
            torch.testing.assert_close(got[k].detach(), v, rtol=1e-5, atol=2e-6, msg=f"{name}/{grp}/{k}")


def test_online_replay_matches_reference_golden():
    """oracle/sampler.py:OnlineReplay vs minibatches sampled by the unmodified reference ReplayBuffer
    (tests/golden/online.npz): ring wrap-around, terminal / time-out episodes, mid-episode sampling."""
    from tests.golden_io import load_online
    from tests.online_script import FIELDS, initial_episodes, replay

    z = load_online()
    for name in [str(c) for c in z["cases"]]:
        maxlen, _, discrete, _ = [int(v) for v in z[f"{name}/cfg"]]
        buf = osampler.OnlineReplay(maxlen, z[f"{name}/script/observations"].shape[1:], bool(discrete))
        for ep in initial_episodes(z, name):
            buf.append_episode(ep.observations, ep.actions, ep.rewards, bool(ep.terminal))
        n = 0
        for j, got, ref in replay(z, name, buf, lambda b, B, f, s: b.sample(B, f, s, 0.99)):
            for k in FIELDS:
                assert got[k].dtype == ref[k].dtype and got[k].shape == ref[k].shape, (name, j, k)
                if k == "rewards":
                    np.testing.assert_allclose(got[k], ref[k], rtol=1e-6, atol=1e-7)
                else:
                    assert np.array_equal(got[k], ref[k]), (name, j, k)
            n += 1
        assert n >= 10


@pytest.mark.parametrize("name", ["awac", "awac_n4"])
def test_awac_oracle_matches_reference_golden(name):
    """AWAC is pinned ahead of its CUDA path: oracle/update.py:AWAC against the unmodified
    reference's metrics and post-step parameters (actor Adam with weight decay, batch-softmax weights, delayed actor)."""
    from tests.golden_io import load_awac

    case = Case(load_awac(), name)
    c = case.cfg
    algo = oupdate.AWAC(int(c["obs"]), int(c["act"]), critics=case.group("init", "q"), policy=case.group("init", "pi"),
                        n_action_samples=int(c["n_action_samples"]),
                        update_actor_interval=int(c["update_actor_interval"]), lam=float(c["lam"]))
    for s in range(case.steps):
        m = algo.update(oupdate.Batch(case.batch(s)), oupdate.Noise(injected=case.noise(s)))
        ref = case.step_metrics(s)
        assert set(m) == set(ref), (s, set(m), set(ref))
        for k, v in ref.items():
            assert abs(m[k] - v) <= 1e-5 * max(1.0, abs(v)), (name, s, k, m[k], v)
    for grp, params in (("q", algo.q), ("pi", algo.pi), ("targ_q", algo.targ_q), ("targ_pi", algo.targ_pi)):
        for k, v in case.group("final", grp).items():
            assert float((params[k].detach() - v).abs().max()) <= 2e-6 * max(1.0, float(v.abs().max())), (name, grp, k)


@pytest.mark.parametrize("name", ["crr", "crr_binary_max_soft"])
def test_crr_oracle_matches_reference_golden(name):
    """CRR, pinned ahead of its CUDA path like AWAC: both weight types, both advantage types, hard and soft targets."""
    from tests.golden_io import load_awac

    case = Case(load_awac(), name)
    c = case.cfg
    algo = oupdate.CRR(int(c["obs"]), int(c["act"]), critics=case.group("init", "q"), policy=case.group("init", "pi"),
                       beta=float(c["beta"]), n_action_samples=int(c["n_action_samples"]),
                       advantage_type="max" if c["adv_max"] else "mean", weight_type="binary" if c["binary"] else "exp",
                       max_weight=float(c["max_weight"]), target_update_type="hard" if c["hard"] else "soft",
                       target_update_interval=int(c["target_update_interval"]))
    for s in range(case.steps):
        m = algo.update(oupdate.Batch(case.batch(s)), oupdate.Noise(injected=case.noise(s)))
        for k, v in case.step_metrics(s).items():
            assert abs(m[k] - v) <= 1e-5 * max(1.0, abs(v)), (name, s, k, m[k], v)
    for grp, params in (("q", algo.q), ("pi", algo.pi), ("targ_q", algo.targ_q), ("targ_pi", algo.targ_pi)):
        for k, v in case.group("final", grp).items():
            assert float((params[k].detach() - v).abs().max()) <= 2e-6 * max(1.0, float(v.abs().max())), (name, grp, k)


def test_plas_oracle_matches_reference_golden():
    """PLAS, pinned ahead of its CUDA path: VAE warm-up steps, then latent-policy TD3 steps with the min/max mix target."""
    from tests.golden_io import load_awac

    case = Case(load_awac(), "plas")
    c = case.cfg
    algo = oupdate.PLAS(int(c["obs"]), int(c["act"]), critics=case.group("init", "q"), policy=case.group("init", "pi"),
                        imitator=case.group("init", "imitator"), warmup_steps=int(c["warmup_steps"]),
                        update_actor_interval=int(c["update_actor_interval"]), lam=float(c["lam"]))
    for s in range(case.steps):
        m = algo.update(oupdate.Batch(case.batch(s)), oupdate.Noise(injected=case.noise(s)))
        ref = case.step_metrics(s)
        assert set(m) == set(ref), (s, set(m), set(ref))
        for k, v in ref.items():
            assert abs(m[k] - v) <= 1e-5 * max(1.0, abs(v)), (s, k, m[k], v)
    for grp, params in (("q", algo.q), ("pi", algo.pi), ("imitator", algo.imitator), ("targ_q", algo.targ_q),
                        ("targ_pi", algo.targ_pi)):
        for k, v in case.group("final", grp).items():
            assert float((params[k].detach() - v).abs().max()) <= 2e-6 * max(1.0, float(v.abs().max())), (grp, k)


@pytest.mark.parametrize("name", ["bear", "bear_gaussian"])
def test_bear_oracle_matches_reference_golden(name):
    """BEAR, pinned ahead of its CUDA path: VAE / temperature / MMD-Lagrange / critic / (warm-up or full) actor steps."""
    from tests.golden_io import load_awac

    case = Case(load_awac(), name)
    c = case.cfg
    algo = oupdate.BEAR(int(c["obs"]), int(c["act"]), critics=case.group("init", "q"), policy=case.group("init", "pi"),
                        imitator=case.group("init", "imitator"), warmup_steps=int(c["warmup_steps"]),
                        n_target_samples=int(c["n_target_samples"]), n_mmd_action_samples=int(c["n_mmd_action_samples"]),
                        mmd_kernel="gaussian" if c["gaussian"] else "laplacian", mmd_sigma=float(c["mmd_sigma"]),
                        lam=float(c["lam"]))
    for s in range(case.steps):
        m = algo.update(oupdate.Batch(case.batch(s)), oupdate.Noise(injected=case.noise(s)))
        ref = case.step_metrics(s)
        assert set(m) == set(ref), (s, set(m), set(ref))
        for k, v in ref.items():
            assert abs(m[k] - v) <= 1e-5 * max(1.0, abs(v)), (name, s, k, m[k], v)
    for grp, params in (("q", algo.q), ("pi", algo.pi), ("imitator", algo.imitator), ("targ_q", algo.targ_q),
                        ("targ_pi", algo.targ_pi), ("log_temp", algo.log_temp), ("log_alpha", algo.log_alpha)):
        for k, v in case.group("final", grp).items():
            assert float((params[k].detach() - v).abs().max()) <= 2e-6 * max(1.0, float(v.abs().max())), (name, grp, k)


def test_discrete_bcq_oracle_matches_reference_golden():
    """DiscreteBCQ, pinned ahead of its CUDA path: Double-DQN target over the imitator-masked greedy action, joint loss."""
    from tests.golden_io import load_awac

    z = load_awac()
    case = Case(z, "discrete_bcq")
    c = case.cfg
    algo = oupdate.DiscreteBCQ((int(c["obs"]),), int(c["act"]), critics=case.group("init", "q"),
                               imitator=case.group("init", "imitator"),
                               target_update_interval=int(c["target_update_interval"]),
                               action_flexibility=float(c["action_flexibility"]), beta=float(c["beta"]))
    for s in range(case.steps):
        m = algo.update(oupdate.Batch(case.batch(s)), None)
        for k, v in case.step_metrics(s).items():
            assert abs(m[k] - v) <= 1e-5 * max(1.0, abs(v)), (s, k, m[k], v)
    for grp, params in (("q", algo.q), ("imitator", algo.imitator), ("targ_q", algo.targ_q)):
        for k, v in case.group("final", grp).items():
            assert float((params[k].detach() - v).abs().max()) <= 2e-6 * max(1.0, float(v.abs().max())), (grp, k)
    with torch.no_grad():
        assert np.array_equal(algo.best_action(torch.tensor(z["discrete_bcq/eval_x"])).numpy(), z["discrete_bcq/predict"])


def test_discrete_sac_oracle_matches_reference_golden():
    """DiscreteSAC, pinned ahead of its CUDA path: temperature, Huber critics on the expectation-form soft target,
    categorical actor, Adam eps = 1e-4, hard target copies."""
    from tests.golden_io import load_awac

    case = Case(load_awac(), "discrete_sac")
    c = case.cfg
    algo = oupdate.DiscreteSAC(int(c["obs"]), int(c["act"]), critics=case.group("init", "q"),
                               policy=case.group("init", "pi"), target_update_interval=int(c["target_update_interval"]))
    for s in range(case.steps):
        m = algo.update(oupdate.Batch(case.batch(s)), None)
        ref = case.step_metrics(s)
        assert set(m) == set(ref)
        for k, v in ref.items():
            assert abs(m[k] - v) <= 1e-5 * max(1.0, abs(v)), (s, k, m[k], v)
    for grp, params in (("q", algo.q), ("pi", algo.pi), ("targ_q", algo.targ_q), ("log_temp", algo.log_temp)):
        for k, v in case.group("final", grp).items():
            assert float((params[k].detach() - v).abs().max()) <= 2e-6 * max(1.0, float(v.abs().max())), (grp, k)


def test_td3_plus_relation_oracle_matches_reference_golden():
    """TD3PlusRelation -- the one algorithm this fork adds to d3rlpy -- pinned ahead of its CUDA path: delayed actor,
    lambda-normalised Q term, B x B relational distillation term, its four extra metrics (stale on critic-only steps)."""
    from tests.golden_io import load_awac

    case = Case(load_awac(), "td3_relation")
    c = case.cfg
    algo = oupdate.TD3PlusRelation(int(c["obs"]), int(c["act"]), critics=case.group("init", "q"),
                                   policy=case.group("init", "pi"))
    for s in range(case.steps):
        m = algo.update(oupdate.Batch(case.batch(s)), oupdate.Noise(injected=case.noise(s)))
        ref = case.step_metrics(s)
        assert set(m) == set(ref), (s, set(m), set(ref))
        for k, v in ref.items():
            assert abs(m[k] - v) <= 1e-5 * max(1.0, abs(v)), (s, k, m[k], v)
    for grp, params in (("q", algo.q), ("pi", algo.pi), ("targ_q", algo.targ_q), ("targ_pi", algo.targ_pi)):
        for k, v in case.group("final", grp).items():
            assert float((params[k].detach() - v).abs().max()) <= 2e-6 * max(1.0, float(v.abs().max())), (grp, k)


@pytest.mark.parametrize("name", ["bc", "discrete_bc"])
def test_bc_oracle_matches_reference_golden(name):
    """BC / DiscreteBC, pinned ahead of their CUDA path: losses, post-step parameters and `predict`."""
    from tests.golden_io import load_awac

    z = load_awac()
    case = Case(z, name)
    c = case.cfg
    algo = oupdate.BC(int(c["obs"]), int(c["act"]), imitator=case.group("init", "imitator"), discrete=bool(c["discrete"]),
                      beta=float(c["beta"]))
    for s in range(case.steps):
        m = algo.update(oupdate.Batch(case.batch(s)), None)
        for k, v in case.step_metrics(s).items():
            assert abs(m[k] - v) <= 1e-5 * max(1.0, abs(v)), (s, k, m[k], v)
    for k, v in case.group("final", "imitator").items():
        assert float((algo.imitator[k].detach() - v).abs().max()) <= 2e-6 * max(1.0, float(v.abs().max())), k
    got = algo.predict(torch.tensor(z[f"{name}/eval_x"])).numpy()
    assert np.array_equal(got, z[f"{name}/predict"]) if c["discrete"] else np.allclose(got, z[f"{name}/predict"], atol=1e-6)
