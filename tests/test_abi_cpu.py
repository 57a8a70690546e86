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
    np.testing.assert_array_equal(ds._meta[:, 3], fr.terminal)
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
