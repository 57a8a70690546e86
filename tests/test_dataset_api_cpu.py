"""CPU: the MDPDataset / Episode surface either side of the sampler -- append, extend, compute_stats, episode returns,
iteration and indexing -- against tests/golden/dataset_api.json, recorded from the unmodified reference's compiled
`d3rlpy.dataset` (tests/golden/make_dataset_api.py).  No device work: the HBM replica is only (re)built on use."""
import json
import os

import numpy as np
import pytest

from tests.golden_io import GOLDEN


def _load():
    with open(os.path.join(GOLDEN, "dataset_api.json")) as f:
        return json.load(f)


def _arrays(lst, discrete):
    o, a, r, t, ept = lst
    return (np.asarray(o, np.float32), np.asarray(a, np.int32 if discrete else np.float32), np.asarray(r, np.float32),
            np.asarray(t, np.float32), np.asarray(ept, np.float32))


def _check(ds, ref, what):
    assert len(ds) == ref["n_episodes"] and ds.size() == ref["size"], what
    assert ds.get_action_size() == ref["action_size"] and list(ds.get_observation_shape()) == ref["observation_shape"]
    assert ds.is_action_discrete() == ref["discrete"]
    assert [e.size() for e in ds.episodes] == ref["episode_sizes"], what
    assert [len(e) for e in ds] == ref["episode_lens"], what                      # __iter__ over episodes
    assert [int(e.observations.shape[0]) for e in ds.episodes] == ref["n_steps_per_episode"]
    assert np.allclose([float(e.compute_return()) for e in ds.episodes], ref["episode_returns"], rtol=1e-6, atol=1e-6)
    assert [float(e.terminal) for e in ds.episodes] == ref["episode_terminal"]
    stats = ds.compute_stats()
    assert set(stats) == set(ref["stats"]), what
    for grp, d in ref["stats"].items():
        assert set(stats[grp]) == set(d), (what, grp)
        for k, v in d.items():
            got = stats[grp][k]
            if k == "histogram":
                if grp == "action" and not ref["discrete"]:
                    for g, r in zip(got, v):
                        assert np.array_equal(g[0], r[0]) and np.allclose(g[1], r[1], rtol=1e-6)
                else:
                    assert np.array_equal(np.asarray(got[0]), np.asarray(v[0])) and np.allclose(got[1], v[1], rtol=1e-6)
            else:
                assert np.allclose(np.asarray(got, np.float64), np.asarray(v), rtol=1e-6, atol=1e-7), (what, grp, k)
    tr = ds[1].transitions                                                         # __getitem__ -> episode
    first, last_next, last_term, first_reward = ref["episode1_first_last"]
    assert np.array_equal(tr[0].observation, np.asarray(first, np.float32))
    assert np.array_equal(tr[-1].next_observation, np.asarray(last_next, np.float32))
    assert float(tr[-1].terminal) == last_term and float(ds.episodes[1][0].reward) == first_reward
    assert [t._t for t in ds.episodes[1]] == [t._t for t in tr]                    # Episode.__iter__


@pytest.mark.parametrize("name", ["continuous", "discrete"])
def test_dataset_append_extend_stats_match_reference(name):
    from d3rlpy_b200.dataset import MDPDataset

    case = _load()[name]
    discrete = name == "discrete"
    ds = MDPDataset(*_arrays(case["inputs"]["base"], discrete), discrete_action=discrete)
    _check(ds, case["base"], "base")
    ds._replays["stale"] = object()
    ds.append(*_arrays(case["inputs"]["more"], discrete))
    assert ds._replays == {}                                   # HBM replicas of the old arrays are dropped
    _check(ds, case["appended"], "appended")
    ds.extend(MDPDataset(*_arrays(case["inputs"]["other"], discrete), discrete_action=discrete))
    _check(ds, case["extended"], "extended")
    with pytest.raises(AssertionError):
        ds.extend(_Mismatch(ds))                               # action type of the other dataset differs
    with pytest.raises(AssertionError):
        bad = _arrays(case["inputs"]["base"], discrete)
        bad[2][3] = np.nan
        MDPDataset(*bad, discrete_action=discrete)


class _Mismatch:
    """A dataset-like object of the other action type."""

    def __init__(self, ds):
        self._ds = ds

    def is_action_discrete(self):
        return not self._ds.is_action_discrete()

    def get_observation_shape(self):
        return self._ds.get_observation_shape()
