"""CPU: bookkeeping of the HBM online replay buffer (append / FIFO window / ring-slot numbering / log compaction and
growth / dirty-range metadata upload) with the device replaced by host tensors and the gather kernels replaced by the
oracle's gather over the SAME logs.  The kernels themselves are covered by tests/test_online_gpu.py."""
import contextlib

import numpy as np
import pytest
import torch

from oracle import sampler as osampler
from tests.golden_io import load_online
from tests.online_script import FIELDS, episode_ns, initial_episodes, replay


class _Stream:
    def __init__(self, device=None):
        pass

    def synchronize(self):
        pass


@pytest.fixture
def host_buffer(monkeypatch):
    import d3rlpy_b200.online.buffers as ob

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())

    def from_indices(view, indices, n_frames=1, n_steps=1, gamma=0.99, scaler=None, out=None):
        n = view.n_transitions
        meta = view.meta[:n].numpy()
        r = osampler.FlatReplay.__new__(osampler.FlatReplay)
        r.observations, r.actions, r.rewards = view.obs.numpy(), view.actions.numpy(), view.rewards.numpy()
        r.discrete = view.discrete
        r.step, r.ep_start, r.ep_last = meta[:, 0], meta[:, 1], meta[:, 2]
        r.terminal, r.zero_next = meta[:, 3] & 1, (meta[:, 3] >> 1) & 1
        assert indices.min() >= 0 and indices.max() < n
        return osampler.gather(r, indices, n_frames, n_steps, gamma)

    monkeypatch.setattr(ob.TransitionMiniBatch, "from_indices", staticmethod(from_indices))
    return lambda *a, **k: ob.ReplayBuffer(*a, device="cpu", **k)


def _same(got, ref, what):
    for k in FIELDS:
        assert got[k].dtype == ref[k].dtype and got[k].shape == ref[k].shape, (what, k)
        if k == "rewards":
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-6, atol=1e-7)
        else:
            assert np.array_equal(got[k], ref[k]), (what, k)


@pytest.mark.parametrize("stage_steps", [4096, 7])
def test_buffer_bookkeeping_reproduces_reference_samples(host_buffer, stage_steps):
    z = load_online()
    for name in [str(c) for c in z["cases"]]:
        buf = host_buffer(int(z[f"{name}/cfg"][0]), episodes=initial_episodes(z, name), stage_steps=stage_steps)
        for j, got, ref in replay(z, name, buf, lambda b, B, f, s: b.sample(B, f, s, 0.99)):
            _same(got, ref, (name, j))


@pytest.mark.parametrize("kind", ["vector", "image"])
def test_buffer_bookkeeping_long_script_with_compaction(host_buffer, kind):
    rs = np.random.RandomState(5)
    maxlen, discrete = 64, kind == "image"
    oshape = (1, 6, 6) if discrete else (7,)
    asize = 4 if discrete else 3
    z0 = np.zeros((2,) + oshape, np.uint8 if discrete else np.float32)
    a0 = np.zeros(2, np.int32) if discrete else np.zeros((2, asize), np.float32)
    buf = host_buffer(maxlen, episodes=[episode_ns(z0, a0, np.zeros(2, np.float32), 0.0, asize)], stage_steps=50)
    orc = osampler.OnlineReplay(maxlen, oshape, discrete)
    orc.append_episode(z0, a0, np.zeros(2, np.float32), False)
    events = checks = 0
    long_done = False
    while events < 3000:
        n = int(rs.randint(1, 30))
        if events >= 900 and not long_done:   # one episode far longer than the buffer and the initial logs
            n, long_done = 400, True
        terminal = rs.rand() < 0.5
        for i in range(n):
            if discrete:
                o, a = rs.randint(0, 256, size=oshape).astype(np.uint8), int(rs.randint(asize))
            else:
                o, a = rs.randn(*oshape).astype(np.float32), rs.uniform(-1, 1, asize).astype(np.float32)
            r, last = float(rs.randn()), i == n - 1
            buf.append(o, a, r, 1.0 if (last and terminal) else 0.0, clip_episode=last)
            orc.append(o, a, r, 1.0 if (last and terminal) else 0.0, clip_episode=last)
            events += 1
            if events % 97 == 0:
                assert len(buf) == len(orc)
                slots = rs.randint(len(buf), size=48)
                for n_frames, n_steps in ((1, 1), (4, 3)) if discrete else ((1, 1), (1, 4)):
                    got = buf.sample_slots(slots, n_frames, n_steps, 0.99)
                    ref = osampler.gather(orc.flat(), orc.transitions_of_slots(slots), n_frames, n_steps, 0.99)
                    _same(got, ref, (kind, events, n_frames, n_steps))
                    checks += 1
    assert checks >= 50 and buf._t_base > 0
    assert buf._cap_s > 3 * maxlen + 64   # the 400-step episode forced the step log to grow
