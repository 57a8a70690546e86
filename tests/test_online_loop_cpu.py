"""CPU: host logic of the online loop and the explorers.  `train_single_env` must call the environment, the buffer and
the algorithm in exactly the order the unmodified reference does (tests/golden/online_loop_trace.json, recorded by
tests/golden/make_online_loop_trace.py with the same recording fakes)."""
import json
import os

import numpy as np
import pytest

from tests.online_loop_fakes import CONFIGS, run


@pytest.mark.parametrize("name", list(CONFIGS))
def test_train_single_env_call_order_matches_reference(name):
    from d3rlpy_b200.online.iterators import train_single_env

    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "online_loop_trace.json")))[name]
    got = json.loads(json.dumps(run(train_single_env, CONFIGS[name])))
    assert got["callback"] == ref["callback"]
    assert len(got["trace"]) == len(ref["trace"])
    for i, (a, b) in enumerate(zip(got["trace"], ref["trace"])):
        assert a == b, (name, i, a, b)


def test_train_single_env_returns_epoch_means():
    from d3rlpy_b200.online.iterators import train_single_env

    cfg = dict(CONFIGS["plain"])
    trace = []
    from tests.online_loop_fakes import FakeAlgo, FakeBuffer, FakeEnv

    hist = train_single_env(FakeAlgo(trace), FakeEnv(trace), FakeBuffer(trace), **cfg)
    assert len(hist) == 4 and all("loss" in h for h in hist)
    # updates start once len(buffer) > batch_size (step 5): epoch 1 averages losses 1..6, epoch 2 losses 7..16
    assert hist[0]["loss"] == pytest.approx(np.mean(np.arange(1, 7)))
    assert hist[1]["loss"] == pytest.approx(np.mean(np.arange(7, 17)))
    assert "rollout_return" in hist[0]


class _Greedy:
    action_size = 5
    action_scaler = None

    def predict(self, x):
        return np.full(x.shape[0], 3) if self.discrete else np.full((x.shape[0], 2), 0.95, np.float32)


def test_explorers_follow_reference_arithmetic_and_numpy_stream():
    """d3rlpy/online/explorers.py:36-171: epsilon-greedy draws `randint` then `random` per call; the linear schedule;
    NormalNoise adds ONE scalar draw to every component and clips to [-1, 1]."""
    from d3rlpy_b200.online import ConstantEpsilonGreedy, LinearDecayEpsilonGreedy, NormalNoise

    algo = _Greedy()
    algo.discrete = True
    x = np.zeros((6, 3), np.float32)
    np.random.seed(3)
    got = ConstantEpsilonGreedy(0.5).sample(algo, x, 0)
    np.random.seed(3)
    rnd = np.random.randint(5, size=6)
    want = np.where(np.random.random(6) < 0.5, rnd, 3)
    assert np.array_equal(got, want)
    lin = LinearDecayEpsilonGreedy(1.0, 0.1, 100)
    assert lin.compute_epsilon(0) == 1.0 and lin.compute_epsilon(100) == 0.1 and lin.compute_epsilon(10 ** 9) == 0.1
    assert lin.compute_epsilon(50) == pytest.approx(0.9 * 0.5 + 0.1)
    np.random.seed(4)
    got = lin.sample(algo, x, 25)
    np.random.seed(4)
    rnd = np.random.randint(5, size=6)
    assert np.array_equal(got, np.where(np.random.random(6) < lin.compute_epsilon(25), rnd, 3))
    algo.discrete = False
    np.random.seed(5)
    got = NormalNoise(0.0, 0.3).sample(algo, x, 0)
    np.random.seed(5)
    assert np.allclose(got, np.clip(0.95 + np.random.normal(0.0, 0.3), -1.0, 1.0)) and got.shape == (6, 2)
    # with a MinMaxActionScaler the clip range is the scaler's [minimum, maximum] (explorers.py:144-151): `predict`
    # returns actions in environment units, which may exceed +-1
    from d3rlpy_b200.preprocessing import MinMaxActionScaler

    algo.action_scaler = MinMaxActionScaler(minimum=np.array([[-2.0, -3.0]]), maximum=np.array([[2.0, 1.0]]))
    algo.predict = lambda x: np.tile(np.array([[1.9, 0.9]], np.float32), (x.shape[0], 1))
    np.random.seed(6)
    got = NormalNoise(0.5, 0.01).sample(algo, x, 0)
    np.random.seed(6)
    n = np.random.normal(0.5, 0.01)
    assert np.allclose(got, np.tile([[min(1.9 + n, 2.0), min(0.9 + n, 1.0)]], (6, 1))) and got[0, 0] > 1.0
