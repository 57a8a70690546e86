"""bench.py's host-side helpers (no GPU): the algorithmic-FLOP model behind `roofline.achieved`, the synthetic data
of BASELINE.md, the reference arm's thread sweep and the minibatches both arms are fed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _mlp_macs(in_dim, hidden, out):
    dims = [in_dim] + list(hidden)
    return sum(a * b for a, b in zip(dims[:-1], dims[1:])) + hidden[-1] * out


def test_required_flops_follow_survey_8d():
    """SURVEY.md 8(d) 'req': c1 ~0.41, c2 19.0, c5 3 497 GFLOP per update.  This model is a little below the survey's
    c2 / c5 figures (the policy trunk runs once on [obs; next_obs] and is shared by all four steps), never above."""
    f = {k: bench.req_gemm_flops(bench.WORKLOADS[k]) / 1e9 for k in ("c1", "c2", "c5")}
    assert abs(f["c1"] - 0.41) < 0.01
    assert 17.5 < f["c2"] <= 19.0
    assert 3300 < f["c5"] <= 3497
    # the dominant term by hand: critic step on R = B (1 + 3N) rows, forward + dgrad + wgrad, E members
    w = bench.WORKLOADS["c2"]
    R = w["batch"] * (1 + 3 * w["n"])
    critic_step = 3 * 2 * R * _mlp_macs(w["obs"] + w["act"], w["hidden"], 1) * w["critics"]
    assert 0.55 < critic_step / (f["c2"] * 1e9) < 0.75


def test_synthetic_dataset_matches_baseline_description():
    w = bench.WORKLOADS["c2"]
    obs, act, rew, term = bench.make_dataset(w, steps_total=5000)
    assert obs.shape == (5000, 17) and act.shape == (5000, 6) and rew.shape == (5000,) and term.shape == (5000,)
    assert obs.dtype == act.dtype == rew.dtype == term.dtype == np.float32
    assert float(np.abs(act).max()) <= 1.0
    assert np.flatnonzero(term).tolist() == [999, 1999, 2999, 3999, 4999]   # episodes of 1 000 steps
    again = bench.make_dataset(w, steps_total=5000)
    assert all(np.array_equal(a, b) for a, b in zip((obs, act, rew, term), again))   # seeded


def test_host_batches_have_the_reference_minibatch_layout():
    w = dict(bench.WORKLOADS["c1"], batch=32)
    obs, act, rew, term = bench.make_dataset(w, steps_total=3000)
    hb = bench.host_batches(w, 3, obs, act, rew, term)
    assert len(hb) == 3
    for b in hb:
        assert b["observations"].shape == (32, 11) and b["next_observations"].shape == (32, 11)
        assert b["actions"].shape == (32, 3)
        for k in ("rewards", "terminals", "n_steps"):
            assert b[k].shape == (32, 1), k
        assert set(np.unique(b["n_steps"])) == {1.0}


def test_thread_sweep_and_config_keys():
    assert bench.thread_sweep(16) == [1, 8, 16] and bench.thread_sweep(1) == [1] and bench.thread_sweep(2) == [1, 2]
    assert bench.CONFIG_KEYS[0] == "workload" and len(set(bench.CONFIG_KEYS)) == len(bench.CONFIG_KEYS)
    assert bench.METRIC == "CQL gradient updates/sec at batch 256"
