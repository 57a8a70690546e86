"""GPU: the HBM online replay buffer (d3rlpy_b200/online/buffers.py) vs minibatches sampled by the unmodified reference
`d3rlpy.online.buffers.ReplayBuffer` under the same numpy seeds (tests/golden/online.npz), and vs the oracle on a long
random script that forces log compaction and growth.  Bit-exact except n-step rewards (allclose 1e-6)."""
import numpy as np
import pytest

from oracle import sampler as osampler
from tests.golden_io import load_online
from tests.online_script import FIELDS, episode_ns, initial_episodes, replay

pytestmark = pytest.mark.gpu


def _check(got, ref, what):
    for k in FIELDS:
        a, b = getattr(got, k), ref[k]
        assert a.dtype == b.dtype and a.shape == b.shape, (what, k, a.dtype, b.dtype, a.shape, b.shape)
        if k == "rewards":
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7, err_msg=str(what))
        else:
            assert np.array_equal(a, b), (what, k)


@pytest.mark.parametrize("stage_steps", [4096, 7])
def test_online_buffer_matches_reference_golden(stage_steps):
    from d3rlpy_b200.online import ReplayBuffer

    z = load_online()
    for name in [str(c) for c in z["cases"]]:
        maxlen = int(z[f"{name}/cfg"][0])
        buf = ReplayBuffer(maxlen, episodes=initial_episodes(z, name), stage_steps=stage_steps)
        n = 0
        for j, got, ref in replay(z, name, buf, lambda b, B, f, s: b.sample(B, f, s, 0.99)):
            _check(got, ref, (name, j))
            n += 1
        assert n >= 10


@pytest.mark.parametrize("kind", ["vector", "image"])
def test_online_buffer_long_script_vs_oracle_with_compaction(kind):
    """3 000 appends into a 64-transition buffer (the device logs hold ~200 rows): many compactions, one episode longer
    than the buffer (its dropped head must stay readable for frame stacks), then an update straight from the sampled
    device batch."""
    from d3rlpy_b200.online import ReplayBuffer

    rs = np.random.RandomState(5)
    maxlen, discrete = 64, kind == "image"
    oshape = (1, 10, 10) if discrete else (7,)
    asize = 4 if discrete else 3

    def step():
        if discrete:
            return rs.randint(0, 256, size=oshape).astype(np.uint8), int(rs.randint(asize)), float(rs.randn())
        return rs.randn(*oshape).astype(np.float32), rs.uniform(-1, 1, asize).astype(np.float32), float(rs.randn())

    z0 = np.zeros((2,) + oshape, np.uint8 if discrete else np.float32)
    a0 = np.zeros(2, np.int32) if discrete else np.zeros((2, asize), np.float32)
    buf = ReplayBuffer(maxlen, episodes=[episode_ns(z0, a0, np.zeros(2, np.float32), 0.0, asize)], stage_steps=50)
    orc = osampler.OnlineReplay(maxlen, oshape, discrete)
    orc.append_episode(z0, a0, np.zeros(2, np.float32), False)
    events, checks, long_done = 0, 0, False
    while events < 3000:
        n = int(rs.randint(1, 30))
        if events >= 900 and not long_done:   # one episode far longer than the buffer and the initial logs
            n, long_done = 400, True
        terminal = rs.rand() < 0.5
        for i in range(n):
            o, a, r = step()
            last = i == n - 1
            buf.append(o, a, r, 1.0 if (last and terminal) else 0.0, clip_episode=last)
            orc.append(o, a, r, 1.0 if (last and terminal) else 0.0, clip_episode=last)
            events += 1
            if events % 97 == 0:
                assert len(buf) == len(orc)
                slots = rs.randint(len(buf), size=48)
                for n_frames, n_steps in ((1, 1), (4, 3)) if discrete else ((1, 1), (1, 4)):
                    got = buf.sample_slots(slots, n_frames, n_steps, 0.99)
                    ref = osampler.gather(orc.flat(), orc.transitions_of_slots(slots), n_frames, n_steps, 0.99)
                    _check(got, ref, (kind, events, n_frames, n_steps))
                    checks += 1
    assert checks >= 50 and buf._t_base > 0   # the logs were compacted
    assert buf._cap_s > 3 * maxlen + 64       # ... and the 400-step episode made the step log grow
    if not discrete:   # the sampled device batch feeds algo.update directly
        from d3rlpy_b200.algos import DDPG

        algo = DDPG(actor_encoder_factory=[32, 32], critic_encoder_factory=[32, 32], batch_size=48)
        algo.create_impl(oshape, asize)
        m = algo.update(buf.sample(48))
        assert np.isfinite(m["critic_loss"]) and np.isfinite(m["actor_loss"])
