"""CPU: the C-ABI library loads and exports every symbol include/d3rlpy_b200.h declares; argument
validation returns error codes (no compute calls, no GPU needed); host-side layout logic."""
import ctypes
import os

import numpy as np
import pytest
import torch

from d3rlpy_b200._lib import LIB_PATH, D3BError, lib, parse_header


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB_PATH), "build with python -m d3rlpy_b200.build"
    dll = ctypes.CDLL(LIB_PATH)
    protos = parse_header()
    assert len(protos) >= 30
    for name in protos:
        assert hasattr(dll, name), name
    assert dll.d3b_abi_version() == 1


def test_error_convention_no_exceptions_across_boundary():
    L = lib()
    with pytest.raises(D3BError, match="n < 0"):
        L.adam_step(None, None, None, None, None, -1, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, None)
    with pytest.raises(D3BError, match="out_features"):
        L.head_forward(None, 0, 0, None, 0, 0, None, 0, None, 0, 0, 4, 64, 8, 1, 0, None)
    with pytest.raises(D3BError, match="empty clip interval or zero divisor"):
        L.scale_rewards(1, 4, 1.0, -1.0, 0.0, 1.0, 1.0, None)          # lo > hi
    with pytest.raises(D3BError, match="empty clip interval or zero divisor"):
        L.scale_rewards(1, 4, -1.0, 1.0, 0.0, 1.0, 0.0, None)          # a degenerate (max == min) reward scaler
    with pytest.raises(D3BError, match="null pointer"):
        L.scale_actions(None, None, None, 4, 3, None)
    with pytest.raises(D3BError, match="bad sizes"):
        L.unscale_actions(None, None, None, 4, 0, None)
    assert L.scale_rewards(None, 0, -1.0, 1.0, 0.0, 1.0, 1.0, None) == 0
    assert L.scale_actions(None, None, None, 0, 3, None) == 0
    # zero-sized work is a no-op, not an error
    assert L.gather_vector(None, 4, None, 2, 0, None, None, None, 0, 1, 0.99, None, None, None, None, None, None, None,
                           None, 0.0, None) == 0


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from d3rlpy_b200.algos import CQL

    with pytest.raises(D3BError, match="no CPU fallback"):
        CQL().create_impl((5,), 2)


def test_transition_meta_matches_oracle_flat_replay():
    from d3rlpy_b200.dataset import MDPDataset
    from oracle.sampler import FlatReplay

    rs = np.random.RandomState(0)
    n = 200
    obs = rs.randn(n, 3).astype(np.float32)
    act = rs.rand(n, 2).astype(np.float32)
    rew = rs.randn(n).astype(np.float32)
    ept = np.zeros(n)
    ept[[9, 10, 57, 120, 199]] = 1   # includes a 1-step episode
    term = ept.copy()
    term[[57, 10]] = 0               # truncated episodes drop their last step
    ds = MDPDataset(obs, act, rew, term, episode_terminals=ept)
    fr = FlatReplay(obs, act, rew, term, ept)
    assert ds._meta.shape[0] == len(fr)
    np.testing.assert_array_equal(ds._meta[:, 0], fr.step)
    np.testing.assert_array_equal(ds._meta[:, 1], fr.ep_start)
    np.testing.assert_array_equal(ds._meta[:, 2], fr.ep_last)
    # flags: bit 0 = terminal, bit 1 = zero next observation; Episode-built data sets both together
    np.testing.assert_array_equal(ds._meta[:, 3] & 1, fr.terminal)
    np.testing.assert_array_equal(ds._meta[:, 3] >> 1, fr.terminal)
    eps = ds.episodes
    assert sum(len(e) for e in eps) == len(fr)
    tr = eps[0].transitions
    assert tr[0].prev_transition is None and tr[-1].next_transition is None and tr[-1].terminal == 1.0


def test_arena_state_dict_keys_match_reference_names():
    from d3rlpy_b200.nets import DenseNet
    from oracle import update as ou

    g = torch.Generator().manual_seed(0)
    q = DenseNet(5 + 2, [8, 8], [("_fc", 1)], 2, torch.device("cpu"), "_encoder.", "_q_funcs.{e}.{name}", True, g)
    assert set(q.arena.state_dict().keys()) == set(ou.make_critics(5, 2, [8, 8], 2, g).keys())
    pi = DenseNet(5, [8, 8, 8], [("_mu", 2), ("_logstd", 2)], 1, torch.device("cpu"), "_encoder.", seed_gen=g)
    assert set(pi.arena.state_dict().keys()) == set(ou.make_squashed_normal_policy(5, 2, [8, 8, 8], g).keys())
    sd = ou.make_squashed_normal_policy(5, 2, [8, 8, 8], g)
    pi.arena.load_state_dict(sd)
    for k, v in pi.arena.state_dict().items():
        assert torch.equal(v, sd[k])
    # heads are physically concatenated: mu rows then logstd rows
    hw = pi.arena.view("__head.weight")
    assert torch.equal(hw[:2], sd["_mu.weight"]) and torch.equal(hw[2:], sd["_logstd.weight"])


def test_fit_index_stream_equals_random_iterator_draws():
    """fit() draws one (chunk, batch) block of indices per epoch; the reference's RandomIterator draws them one by
    one with np.random.randint(n) (iterators/random_iterator.py:38-41).  Same legacy stream => same indices."""
    n, B, steps = 987_654, 256, 7
    a, b = np.random.RandomState(11), np.random.RandomState(11)
    seq = np.array([[a.randint(n) for _ in range(B)] for _ in range(steps)])
    blk = b.randint(n, size=(steps, B))
    assert np.array_equal(seq, blk)


def test_header_prototypes_parse_and_cover_all_sources():
    """Every extern "C" d3b_* definition in csrc/ is declared in include/d3rlpy_b200.h (and vice versa)."""
    import glob
    import re

    here = os.path.dirname(os.path.abspath(__file__))
    declared = set(parse_header().keys())
    defined = set()
    for f in glob.glob(os.path.join(here, "..", "d3rlpy_b200", "csrc", "*.cu")):
        defined |= set(re.findall(r'extern "C"\s+[\w\s\*]+?\b(d3b_\w+)\s*\(', open(f).read()))
    assert declared == defined, (sorted(declared - defined), sorted(defined - declared))


def test_params_json_matches_reference_format():
    """`save_params` (d3rlpy/base.py:823-850): our document holds every key the unmodified reference writes, with the
    same value encodings (tests/golden/params_json.json, generated by tests/golden/make_params_json.py); the only
    extra keys are this package's own constructor arguments."""
    import json
    import os
    from types import SimpleNamespace

    from d3rlpy_b200.algos import (AWAC, BCQ, BEAR, CQL, CRR, DDPG, DQN, IQL, NFQ, PLAS, SAC, TD3, DiscreteCQL,
                                   QRQFunctionFactory, TD3PlusBC)

    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "params_json.json")))
    enc = [32, 32]
    cases = {
        "ddpg": (DDPG(actor_encoder_factory=enc, critic_encoder_factory=enc, use_gpu=None), (6,), 3),
        "dqn_qr": (DQN(encoder_factory=enc, q_func_factory="qr", use_gpu=None), (6,), 4),
        "nfq": (NFQ(encoder_factory=enc, use_gpu=None), (6,), 4),
        "iql": (IQL(actor_encoder_factory=enc, critic_encoder_factory=enc, value_encoder_factory=enc, use_gpu=None),
                (6,), 3),
        "dcql_qr": (DiscreteCQL(encoder_factory=enc, q_func_factory=QRQFunctionFactory(n_quantiles=16), use_gpu=None),
                    (6,), 4),
        "cql": (CQL(actor_encoder_factory=enc, critic_encoder_factory=enc, n_action_samples=4, use_gpu=None), (6,), 3),
        "td3bc": (TD3PlusBC(actor_encoder_factory=enc, critic_encoder_factory=enc, use_gpu=None), (6,), 3),
        "bcq": (BCQ(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=enc, use_gpu=None),
                (6,), 3),
        "dcql": (DiscreteCQL(encoder_factory=enc, n_critics=2, use_gpu=None), (6,), 4),
        "dcql_pixel": (DiscreteCQL(n_frames=4, scaler="pixel", use_gpu=None), (4, 84, 84), 4),
        "sac": (SAC(actor_encoder_factory=enc, critic_encoder_factory=enc, use_gpu=None), (6,), 3),
        "td3": (TD3(actor_encoder_factory=enc, critic_encoder_factory=enc, use_gpu=None), (6,), 3),
        "awac": (AWAC(actor_encoder_factory=enc, critic_encoder_factory=enc, n_action_samples=2, use_gpu=None), (6,), 3),
        "crr": (CRR(actor_encoder_factory=enc, critic_encoder_factory=enc, advantage_type="max", weight_type="binary",
                    use_gpu=None), (6,), 3),
        "plas": (PLAS(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=enc, lam=0.6,
                      use_gpu=None), (6,), 3),
        "bear": (BEAR(actor_encoder_factory=enc, critic_encoder_factory=enc, imitator_encoder_factory=enc,
                      mmd_kernel="gaussian", n_mmd_action_samples=3, use_gpu=None), (6,), 3),
    }
    for name, (algo, obs, act) in cases.items():
        algo._impl = SimpleNamespace(observation_shape=obs, action_size=act)  # the document needs the shapes only
        doc = json.loads(json.dumps(algo._params_document()))
        for key, value in ref[name].items():
            assert key in doc, (name, key)
            assert doc[key] == value, (name, key, doc[key], value)
        assert set(doc) - set(ref[name]) <= {"seed", "precision"}, (name, set(doc) - set(ref[name]))
        # the three shape/identity keys come last, as in the reference's file
        assert list(doc)[-3:] == ["algorithm", "observation_shape", "action_size"]


def test_round_iterator_index_stream_matches_reference_walk():
    """`fit(n_epochs=...)` cuts one shuffled permutation per epoch into consecutive batches; the reference's
    RoundIterator (iterators/round_iterator.py:39-55 + base.py:53 `[self.get_next() for _ in range(batch_size)]`)
    walks `self._indices[self._index]` one transition at a time after `np.random.shuffle`, for
    `len(transitions) // batch_size` batches per epoch.  Same legacy stream => same batches; remainder dropped."""
    from d3rlpy_b200.algos.base import round_iterator_indices

    n, B = 1003, 32
    a, b = np.random.RandomState(5), np.random.RandomState(5)
    for _ in range(3):  # consecutive epochs keep consuming the same stream
        indices = np.arange(n)
        a.shuffle(indices)
        index, walk = 0, []
        for _ in range(n // B):
            walk.append([int(indices[index + j]) for j in range(B)])
            index += B
        got = round_iterator_indices(b, n, B, shuffle=True)
        assert got.shape == (n // B, B) and np.array_equal(got, np.array(walk))
    plain = round_iterator_indices(b, n, B, shuffle=False)
    assert np.array_equal(plain.ravel(), np.arange(n // B * B))


def test_exported_greedy_policies_match_oracle(tmp_path):
    """d3rlpy_b200/export.py (`save_policy`, algos/torch/base.py:86-126): the traced TorchScript functions rebuilt from
    state_dicts in the reference key layout give the oracle's greedy actions (scaler included); file round trip."""
    import torch

    from d3rlpy_b200.export import GreedyPolicy, save_policy
    from d3rlpy_b200.preprocessing import StandardScaler
    from oracle import update as ou

    rs = np.random.RandomState(0)
    O, A, n = 9, 4, 11
    x = torch.tensor(rs.randn(n, O).astype(np.float32))
    # CQL: tanh(mu)
    orc = ou.CQL(O, A, hidden=[32, 32], n_action_samples=3, seed=1)
    f = str(tmp_path / "cql.pt")
    save_policy(GreedyPolicy("normal", policy=orc.pi, q=orc.q), (O,), f)
    with torch.no_grad():
        np.testing.assert_allclose(torch.jit.load(f)(x).numpy(), ou.policy_best_action(orc.pi, x).numpy(), rtol=1e-6, atol=1e-6)
    # TD3+BC with a fitted standard scaler
    sc = StandardScaler(mean=x.mean(0).numpy(), std=x.std(0).numpy())
    orc = ou.TD3PlusBC(O, A, hidden=[32, 32], seed=2)
    f = str(tmp_path / "td3.pt")
    save_policy(GreedyPolicy("deterministic", policy=orc.pi, q=orc.q, scaler=sc), (O,), f)
    xs = ou.standard_scaler(x.mean(0).numpy(), x.std(0).numpy())(x)
    with torch.no_grad():
        np.testing.assert_allclose(torch.jit.load(f)(x).numpy(), ou.deterministic_policy(orc.pi, xs).numpy(), rtol=1e-5, atol=1e-6)
    # DiscreteCQL on pixels: argmax of the member-mean Q of x / 255
    orc = ou.DiscreteCQL((4, 84, 84), 4, n_critics=2, seed=3)
    frames = torch.tensor(rs.randint(0, 256, (3, 4, 84, 84)).astype(np.float32))
    f = str(tmp_path / "dcql.pt")
    save_policy(GreedyPolicy("discrete", q=orc.q, scaler="pixel"), (4, 84, 84), f)
    with torch.no_grad():
        ref = ou.q_discrete(orc.q, frames / 255.0, "mean").argmax(1)
        assert torch.equal(torch.jit.load(f)(frames), ref)
    # BCQ: candidates are sampled inside the traced function -> check range / shape and that the choice is a candidate
    orc = ou.BCQ(O, A, hidden=[32, 32], vae_hidden=[48, 48], n_action_samples=5, seed=4)
    f = str(tmp_path / "bcq.pt")
    save_policy(GreedyPolicy("bcq", policy=orc.pi, q=orc.q, imitator=orc.imitator, n_action_samples=5), (O,), f)
    with torch.no_grad():
        out = torch.jit.load(f)(x)
    assert out.shape == (n, A) and float(out.abs().max()) <= 1.0
    with pytest.raises(ValueError):
        save_policy(GreedyPolicy("normal", policy=ou.CQL(O, A, hidden=[32, 32], seed=1).pi), (O,), str(tmp_path / "p.bin"))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU reference arm: the oracle port timed on the host cores, no GPU needed):
    ONE JSON line with the arm's keys — same metric / unit / config as our arm, `impl`, `cpu_baseline` describing the
    run and an `e2e` object without copies."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "CQL gradient updates/sec at batch 256" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("CQL halfcheetah-shaped")
    sys.path.insert(0, root)
    try:
        import bench
    finally:
        sys.path.pop(0)
    assert tuple(d["config"]) == bench.CONFIG_KEYS   # the GPU arm asserts the same key set on its own line
    assert d["warmup"] == 3 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
