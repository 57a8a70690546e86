"""GPU: the replay gather kernels (K1/K1b) vs the oracle sampler and the reference's golden vectors.
Bit-exact for observations / frame stacks / actions / terminals / n_steps; rewards allclose(1e-6)
for n_steps > 1 (double-pow accumulate under -ffast-math in the reference)."""
import numpy as np
import pytest
import torch

from oracle import sampler as osampler
from tests.golden_io import load_sampler

pytestmark = pytest.mark.gpu


def _check(out, ref, name=""):
    for k in ("observations", "next_observations", "actions", "terminals", "n_steps"):
        a, b = getattr(out, k), ref[k]
        assert a.dtype == b.dtype and a.shape == b.shape, (name, k, a.dtype, b.dtype, a.shape, b.shape)
        assert np.array_equal(a, b), (name, k)
    np.testing.assert_allclose(out.rewards, ref["rewards"], rtol=1e-6, atol=1e-7)


def test_gather_matches_reference_golden_all_cases():
    from d3rlpy_b200.dataset import MDPDataset, TransitionMiniBatch

    z = load_sampler()
    for name in [str(c) for c in z["cases"]]:
        d = lambda k: z[f"{name}/data/{k}"]
        ds = MDPDataset(d("observations"), d("actions"), d("rewards"), d("terminals"), d("episode_terminals"),
                        discrete_action="disc" in name)
        n_frames, n_steps = [int(v) for v in z[f"{name}/cfg"]]
        trs = ds.transitions()
        batch = TransitionMiniBatch([trs[i] for i in z[f"{name}/indices"]], n_frames=n_frames, n_steps=n_steps,
                                    gamma=0.99)
        ref = {k: z[f"{name}/ref/{k}"] for k in ("observations", "next_observations", "actions", "terminals",
                                                 "n_steps", "rewards")}
        _check(batch, ref, name)


@pytest.mark.parametrize("kind", ["vector", "atari"])
def test_gather_full_size_vs_oracle_and_properties(kind):
    """BASELINE-sized shapes: c5-like vector rows (B=8192, O=111) and Atari 84x84 x4 stacks (B=32)."""
    from d3rlpy_b200.dataset import MDPDataset, TransitionMiniBatch

    rs = np.random.RandomState(0)
    if kind == "vector":
        S, O, A, B, ep = 60_000, 111, 8, 8192, 1000
        obs = rs.randn(S, O).astype(np.float32)
        act = rs.uniform(-1, 1, (S, A)).astype(np.float32)
        n_frames, n_steps = 1, 3
    else:
        S, B, ep = 6000, 32, 2000
        obs = rs.randint(0, 256, (S, 1, 84, 84)).astype(np.uint8)
        act = rs.randint(0, 4, S).astype(np.int32)
        n_frames, n_steps = 4, 1
    rew = rs.randn(S).astype(np.float32)
    term = np.zeros(S, np.float32)
    term[ep - 1::ep] = 1
    ds = MDPDataset(obs, act, rew, term, discrete_action=(kind == "atari"))
    replay = ds.device_replay("cuda:0")
    idx = rs.randint(len(replay), size=B)
    idx[:3] = [0, ep - 1, ep]  # episode start / terminal / next episode start
    batch = TransitionMiniBatch.from_indices(replay, idx, n_frames=n_frames, n_steps=n_steps, gamma=0.99)
    ref = osampler.gather(osampler.FlatReplay(obs, act, rew, term), idx, n_frames, n_steps, 0.99)
    _check(batch, ref, kind)
    # size-independent properties: terminal rows have all-zero next observations; n_steps in range;
    # idempotence (same indices -> same bytes)
    nxt, t = batch.next_observations, batch.terminals.reshape(-1)
    assert not nxt[t == 1].any()
    assert batch.n_steps.min() >= 1 and batch.n_steps.max() <= n_steps
    again = TransitionMiniBatch.from_indices(replay, idx, n_frames=n_frames, n_steps=n_steps, gamma=0.99)
    assert np.array_equal(again.observations, batch.observations)
    if kind == "atari":  # frame stacking: newest channel of obs equals the raw frame
        assert np.array_equal(batch.observations[:, -1], obs[ds._meta[idx, 0], 0])
        assert np.array_equal(batch.observations[0, 0], batch.observations[0, 3])  # idx 0: start padding


def test_gather_fused_standard_scaler():
    from d3rlpy_b200.dataset import MDPDataset, TransitionMiniBatch
    from d3rlpy_b200.preprocessing import StandardScaler

    rs = np.random.RandomState(1)
    S, O, A, B = 5000, 11, 3, 256
    obs = (rs.randn(S, O) * 3 + 1).astype(np.float32)
    ds = MDPDataset(obs, rs.uniform(-1, 1, (S, A)).astype(np.float32), rs.randn(S), (np.arange(S) % 500 == 499))
    sc = StandardScaler(ds)
    replay = ds.device_replay("cuda:0")
    idx = rs.randint(len(replay), size=B)
    raw = TransitionMiniBatch.from_indices(replay, idx)
    scaled = TransitionMiniBatch.from_indices(replay, idx, scaler=sc)
    mean = torch.tensor(sc._mean, dtype=torch.float32)
    std = torch.tensor(sc._std, dtype=torch.float32)
    expect = (torch.tensor(raw.observations) - mean) / (std + 1e-3)   # scalers.py:350-354
    assert torch.equal(torch.tensor(scaled.observations), expect)


def test_gather_empty_batch_is_noop():
    from d3rlpy_b200._lib import lib

    assert lib().gather_frames(None, 16, None, None, 0, 4, 1, None, None, None) == 0


@pytest.mark.parametrize("discrete", [False, True])
def test_gather_vector_large_ragged_batch(discrete):
    """20 011 rows (not a multiple of the 8 rows per block), n_steps = 3, episodes of random length ending in terminals
    or time-outs, continuous and discrete actions, vs the oracle."""
    from d3rlpy_b200.dataset import MDPDataset, TransitionMiniBatch

    rs = np.random.RandomState(11)
    S, O, A, B = 6000, 13, 2, 20_011
    obs = rs.randn(S, O).astype(np.float32)
    act = rs.randint(0, 5, S).astype(np.int32) if discrete else rs.uniform(-1, 1, (S, A)).astype(np.float32)
    rew = rs.randn(S).astype(np.float32)
    ends = np.zeros(S, np.float32)
    ends[np.cumsum(rs.randint(2, 40, size=400))[:-1].clip(max=S - 1)] = 1
    ends[-1] = 1
    term = ends * (rs.rand(S) < 0.5)
    ds = MDPDataset(obs, act, rew, term, ends, discrete_action=discrete)
    replay = ds.device_replay("cuda:0")
    idx = rs.randint(len(replay), size=B)
    batch = TransitionMiniBatch.from_indices(replay, idx, n_steps=3, gamma=0.99)
    ref = osampler.gather(osampler.FlatReplay(obs, act, rew, term, ends), idx, 1, 3, 0.99)
    _check(batch, ref, "large ragged batch")
