"""GPU: AWAC, CRR, PLAS and BEAR (`algo.update(batch)` through the public API) against the golden vectors recorded from the
unmodified reference (tests/golden/update_awac.npz, make_golden_awac.py): identical weights, minibatches and injected
noise; metrics per step and post-update parameters (incl. target networks and the logstd parameter).
Tolerance: 1e-5 relative in fp32 mode (eager and graph), bf16 mode: metrics 1e-2, update relative L2 <= 0.3."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests.golden_io import Case, load_awac
from tests.test_update_gpu import BF16_REL, BF16_UPDATE_REL, REL, _assert_metrics, _assert_params, _assert_update

pytestmark = pytest.mark.gpu


def _ns(arrays):
    return SimpleNamespace(**arrays)


def _load(impl, case):
    impl.q_function.load_state_dict(case.group("init", "q"))
    impl.targ_q_function.load_state_dict(case.group("init", "q"))
    impl.policy.load_state_dict(case.group("init", "pi"))
    impl.targ_policy.load_state_dict(case.group("init", "pi"))


def _check_final(impl, case, precision):
    for grp, view in (("q", impl.q_function), ("pi", impl.policy), ("targ_q", impl.targ_q_function),
                      ("targ_pi", impl.targ_policy)):
        if precision == "fp32":
            _assert_params(view.state_dict(), case.group("final", grp), grp, rel=REL)
        if grp in ("q", "pi"):
            _assert_update(view.state_dict(), case.group("final", grp), case.group("init", grp), grp,
                           1e-3 if precision == "fp32" else BF16_UPDATE_REL)


@pytest.mark.parametrize("name,precision,use_graph", [("awac", "fp32", False), ("awac", "fp32", True),
                                                      ("awac_n4", "fp32", True), ("awac_n4", "bf16", True)])
def test_awac_matches_reference_golden(name, precision, use_graph):
    from d3rlpy_b200.algos import AWAC

    case = Case(load_awac(), name)
    c = case.cfg
    B = int(c["batch"])
    algo = AWAC(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=B,
                n_action_samples=int(c["n_action_samples"]), update_actor_interval=int(c["update_actor_interval"]),
                lam=float(c["lam"]), precision=precision)
    algo.create_impl((int(c["obs"]),), int(c["act"]))
    impl = algo.impl
    impl.use_graph = use_graph
    _load(impl, case)
    assert list(impl.policy.state_dict().keys())[0] == "_logstd"   # the parameter precedes the sub-modules
    rel = REL if precision == "fp32" else BF16_REL
    for s in range(case.steps):
        noise = case.noise(s)   # the weights draw only exists on actor steps
        impl.inject_noise(noise, B, names=["target", "weights"][:len(noise)])
        m = algo.update(_ns(case.batch(s)))
        _assert_metrics(m, case.step_metrics(s), f"{name} {precision} step {s}", rel=rel)
    _check_final(impl, case, precision)
    assert algo.grad_step == case.steps


@pytest.mark.parametrize("name,precision,use_graph", [("crr", "fp32", False), ("crr", "fp32", True),
                                                      ("crr_binary_max_soft", "fp32", True), ("crr", "bf16", True)])
def test_crr_matches_reference_golden(name, precision, use_graph):
    from d3rlpy_b200.algos import CRR

    case = Case(load_awac(), name)
    c = case.cfg
    B = int(c["batch"])
    algo = CRR(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=B, beta=float(c["beta"]),
               n_action_samples=int(c["n_action_samples"]), advantage_type="max" if c["adv_max"] else "mean",
               weight_type="binary" if c["binary"] else "exp", max_weight=float(c["max_weight"]),
               target_update_type="hard" if c["hard"] else "soft", target_update_interval=int(c["target_update_interval"]),
               n_critics=len({k.split(".")[1] for k in case.group("init", "q")}), precision=precision)
    algo.create_impl((int(c["obs"]),), int(c["act"]))
    impl = algo.impl
    impl.use_graph = use_graph
    _load(impl, case)
    rel = REL if precision == "fp32" else BF16_REL
    for s in range(case.steps):
        impl.inject_noise(case.noise(s), B)
        m = algo.update(_ns(case.batch(s)))
        _assert_metrics(m, case.step_metrics(s), f"{name} {precision} step {s}", rel=rel)
    _check_final(impl, case, precision)


def test_awac_hooks_and_predict():
    """update_critic / update_actor / compute_actor_loss on their own, predict / sample_action shapes."""
    from d3rlpy_b200.algos import AWAC

    case = Case(load_awac(), "awac")
    c = case.cfg
    B = int(c["batch"])
    algo = AWAC(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=B)
    algo.create_impl((int(c["obs"]),), int(c["act"]))
    impl = algo.impl
    _load(impl, case)
    ref = case.step_metrics(0)
    impl.inject_noise(case.noise(0), B)
    b = _ns(case.batch(0))
    loss0 = float(impl.compute_actor_loss(b))          # before the critic step: finite, nothing stepped
    assert np.isfinite(loss0)
    c_loss = float(impl.update_critic(b))
    assert abs(c_loss - ref["critic_loss"]) <= REL * max(1.0, abs(ref["critic_loss"]))
    a_loss, mean_std = impl.update_actor(b)
    assert abs(float(a_loss) - ref["actor_loss"]) <= REL * max(1.0, abs(ref["actor_loss"]))
    assert abs(float(mean_std) - ref["mean_std"]) <= REL
    x = np.asarray(case.batch(0)["observations"])
    assert algo.predict(x).shape == (B, int(c["act"])) and algo.sample_action(x).shape == (B, int(c["act"]))
    v = algo.predict_value(x, np.asarray(case.batch(0)["actions"]))
    assert v.shape == (B,)


@pytest.mark.parametrize("precision,use_graph", [("fp32", False), ("fp32", True), ("bf16", True)])
def test_plas_matches_reference_golden(precision, use_graph):
    """PLAS: two VAE warm-up steps, then critic every step and the latent-policy actor step every other step
    (tests/golden/update_awac.npz "plas", recorded from the unmodified reference)."""
    from d3rlpy_b200.algos import PLAS

    case = Case(load_awac(), "plas")
    c = case.cfg
    B = int(c["batch"])
    algo = PLAS(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], imitator_encoder_factory=[48, 48],
                batch_size=B, warmup_steps=int(c["warmup_steps"]), update_actor_interval=int(c["update_actor_interval"]),
                lam=float(c["lam"]), precision=precision)
    algo.create_impl((int(c["obs"]),), int(c["act"]))
    impl = algo.impl
    impl.use_graph = use_graph
    _load(impl, case)
    impl.imitator.load_state_dict(case.group("init", "imitator"))
    rel = REL if precision == "fp32" else BF16_REL
    for s in range(case.steps):
        noise = case.noise(s)
        if noise:
            impl.inject_noise(noise, B)
        m = algo.update(_ns(case.batch(s)))
        _assert_metrics(m, case.step_metrics(s), f"plas {precision} step {s}", rel=rel)
    for grp, view in (("q", impl.q_function), ("pi", impl.policy), ("imitator", impl.imitator),
                      ("targ_q", impl.targ_q_function), ("targ_pi", impl.targ_policy)):
        if precision == "fp32":
            _assert_params(view.state_dict(), case.group("final", grp), grp, rel=REL)
        if grp in ("q", "pi", "imitator"):
            _assert_update(view.state_dict(), case.group("final", grp), case.group("init", grp), grp,
                           1e-3 if precision == "fp32" else BF16_UPDATE_REL)
    x = np.asarray(case.batch(0)["observations"])
    assert algo.predict(x).shape == (B, int(c["act"]))


_BEAR_NOISE = ["imitator", "temp", "mmd_lat_alpha", "mmd_eps_alpha", "target", "actor", "mmd_lat_actor", "mmd_eps_actor"]


def _bear(case, precision):
    from d3rlpy_b200.algos import BEAR

    c = case.cfg
    algo = BEAR(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], imitator_encoder_factory=[48, 48],
                batch_size=int(c["batch"]), warmup_steps=int(c["warmup_steps"]),
                n_target_samples=int(c["n_target_samples"]), n_mmd_action_samples=int(c["n_mmd_action_samples"]),
                mmd_kernel="gaussian" if c["gaussian"] else "laplacian", mmd_sigma=float(c["mmd_sigma"]),
                lam=float(c["lam"]), precision=precision)
    algo.create_impl((int(c["obs"]),), int(c["act"]))
    _load(algo.impl, case)
    algo.impl.imitator.load_state_dict(case.group("init", "imitator"))
    return algo


def _bear_noise_names(noise):
    # the SAC actor draw only exists after warm-up (7 draws during warm-up, 8 afterwards)
    return _BEAR_NOISE if len(noise) == 8 else [n for n in _BEAR_NOISE if n != "actor"]


@pytest.mark.parametrize("name,precision,use_graph", [("bear", "fp32", False), ("bear", "fp32", True),
                                                      ("bear_gaussian", "fp32", True), ("bear", "bf16", True)])
def test_bear_matches_reference_golden(name, precision, use_graph):
    """BEAR: two warm-up steps (actor on the MMD loss alone), then SAC + MMD actor steps; Laplacian and Gaussian
    kernels (tests/golden/update_awac.npz "bear" / "bear_gaussian", recorded from the unmodified reference)."""
    case = Case(load_awac(), name)
    B = int(case.cfg["batch"])
    algo = _bear(case, precision)
    impl = algo.impl
    impl.use_graph = use_graph
    rel = REL if precision == "fp32" else BF16_REL
    for s in range(case.steps):
        noise = case.noise(s)
        impl.inject_noise(noise, B, names=_bear_noise_names(noise))
        m = algo.update(_ns(case.batch(s)))
        _assert_metrics(m, case.step_metrics(s), f"{name} {precision} step {s}", rel=rel)
    for grp, view in (("q", impl.q_function), ("pi", impl.policy), ("imitator", impl.imitator),
                      ("targ_q", impl.targ_q_function), ("targ_pi", impl.targ_policy), ("log_temp", impl._log_temp),
                      ("log_alpha", impl._log_alpha)):
        if precision == "fp32":
            _assert_params(view.state_dict(), case.group("final", grp), grp, rel=REL)
        if grp in ("q", "pi", "imitator"):
            _assert_update(view.state_dict(), case.group("final", grp), case.group("init", grp), grp,
                           1e-3 if precision == "fp32" else BF16_UPDATE_REL)
    x = np.asarray(case.batch(0)["observations"])
    assert algo.predict(x).shape == (B, int(case.cfg["act"]))


def test_bear_hooks():
    """The reference hooks one by one reproduce the first golden step's metrics."""
    case = Case(load_awac(), "bear")
    B = int(case.cfg["batch"])
    algo = _bear(case, "fp32")
    impl = algo.impl
    ref, b, noise = case.step_metrics(0), _ns(case.batch(0)), case.noise(0)
    close = lambda got, key: abs(float(got) - ref[key]) <= REL * max(1.0, abs(ref[key]))
    impl.inject_noise(noise, B, names=_bear_noise_names(noise))
    assert close(impl.update_imitator(b), "imitator_loss")
    t_loss, temp = impl.update_temp(b)
    assert close(t_loss, "temp_loss") and close(temp, "temp")
    a_loss, alpha = impl.update_alpha(b)
    assert close(a_loss, "alpha_loss") and close(alpha, "alpha")
    assert impl.compute_target(b).shape == (B, 1)
    assert close(impl.update_critic(b), "critic_loss")
    assert close(impl.warmup_actor(b), "actor_loss")
    with pytest.raises(ValueError):
        from d3rlpy_b200.algos import BEAR

        BEAR(mmd_kernel="cauchy").create_impl((6,), 3)


@pytest.mark.parametrize("name", ["awac", "crr", "plas", "bear"])
def test_f4_checkpoint_and_params_round_trip(name, tmp_path):
    """save_model / load_model and save_params / from_json for the four siblings: a fresh algorithm restored from the
    files continues with bit-identical metrics (Adam moments, step counters, targets and scalars included)."""
    import d3rlpy_b200.algos as algos

    cls = {"awac": algos.AWAC, "crr": algos.CRR, "plas": algos.PLAS, "bear": algos.BEAR}[name]
    O, A, B = 6, 3, 16
    kw = dict(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=B)
    if name in ("plas", "bear"):
        kw["imitator_encoder_factory"] = [48, 48]
        kw["warmup_steps"] = 1
    algo = cls(**kw)
    algo.create_impl((O,), A)
    rs = np.random.RandomState(3)

    def batch():
        return _ns(dict(observations=rs.randn(B, O).astype(np.float32),
                        actions=rs.uniform(-1, 1, (B, A)).astype(np.float32), rewards=rs.randn(B, 1).astype(np.float32),
                        next_observations=rs.randn(B, O).astype(np.float32), terminals=np.zeros((B, 1), np.float32),
                        n_steps=np.ones((B, 1), np.float32)))

    for _ in range(3):
        algo.update(batch())
    f, pj = str(tmp_path / "model.pt"), str(tmp_path / "params.json")
    algo.impl.save_model(f)
    algo.save_params(pj)
    other = cls.from_json(pj)
    assert {k: v for k, v in other.get_params().items() if "factory" not in k} == \
        {k: v for k, v in algo.get_params().items() if "factory" not in k}
    other.impl.load_model(f)
    other.set_grad_step(algo.grad_step)
    other.impl._counters.copy_(algo.impl._counters)
    nb = batch()
    m1, m2 = algo.update(nb), other.update(nb)
    assert m1.keys() == m2.keys()
    for k in m1:
        assert float(m1[k]) == float(m2[k]), (name, k, float(m1[k]), float(m2[k]))


@pytest.mark.parametrize("name", ["awac", "crr", "plas", "bear"])
def test_f4_from_json_reads_reference_params(name, tmp_path):
    """`from_json` on the params.json the unmodified reference writes for these algorithms (tests/golden/params_json.json)."""
    import json
    import os

    import d3rlpy_b200.algos as algos

    cls = {"awac": algos.AWAC, "crr": algos.CRR, "plas": algos.PLAS, "bear": algos.BEAR}[name]
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "params_json.json")))[name]
    f = tmp_path / "params.json"
    f.write_text(json.dumps(ref))
    algo = cls.from_json(str(f), use_gpu=0)
    assert algo.impl is not None and algo.impl.observation_shape == (6,) and algo.impl.action_size == 3
    assert algo._actor_hidden == [32, 32] and algo._critic_hidden == [32, 32]
    if name == "awac":
        assert algo._actor_weight_decay == 1e-4 and algo._n_action_samples == 2
    if name == "crr":
        assert algo._advantage_type == "max" and algo._weight_type == "binary"
    if name == "plas":
        assert algo._lam == 0.6
    if name == "bear":
        assert algo._mmd_kernel == "gaussian" and algo._n_mmd_action_samples == 3
    doc = algo._params_document()
    for key, value in ref.items():
        if key != "use_gpu":
            assert json.loads(json.dumps(doc[key])) == value, (key, doc[key], value)


@pytest.mark.parametrize("name", ["awac", "crr", "plas", "bear"])
def test_f4_fit_loop_and_evaluation_api(name):
    """`fit()` over an HBM-resident dataset (device gather + update graph per step), then the evaluation calls the
    reference's scorers use: shapes, finiteness, action range."""
    import d3rlpy_b200.algos as algos
    from d3rlpy_b200.dataset import MDPDataset

    cls = {"awac": algos.AWAC, "crr": algos.CRR, "plas": algos.PLAS, "bear": algos.BEAR}[name]
    rs = np.random.RandomState(0)
    S, O, A = 4000, 6, 3
    ds = MDPDataset(rs.randn(S, O).astype(np.float32), rs.uniform(-1, 1, (S, A)).astype(np.float32),
                    rs.randn(S).astype(np.float32), (np.arange(S) % 200 == 199).astype(np.float32))
    kw = dict(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=64)
    if name in ("plas", "bear"):
        kw.update(imitator_encoder_factory=[48, 48], warmup_steps=5)
    algo = cls(**kw)
    hist = algo.fit(ds, n_steps=20, n_steps_per_epoch=10, seed=0)
    assert len(hist) == 2 and algo.grad_step == 20
    for _, metrics in hist:
        assert "critic_loss" in metrics or name in ("plas",) and "imitator_loss" in metrics
        assert all(np.isfinite(float(v)) for v in metrics.values()), metrics
    x = rs.randn(9, O).astype(np.float32)
    act = algo.predict(x)
    assert act.shape == (9, A) and np.all(np.isfinite(act)) and np.all(np.abs(act) <= 1.0 + 1e-6)
    v = algo.predict_value(x, act)
    assert v.shape == (9,) and np.all(np.isfinite(v))
    if name != "bear":
        s = algo.sample_action(x)
        assert s.shape == (9, A) and np.all(np.isfinite(s))
