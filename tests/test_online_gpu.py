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


class _Box:
    def __init__(self, shape, rs):
        self.shape, self._rs = shape, rs

    def sample(self):
        return self._rs.uniform(-1, 1, self.shape).astype(np.float32)


class _Discrete:
    def __init__(self, n, rs):
        self.n, self._rs = n, rs

    def sample(self):
        return int(self._rs.randint(self.n))


class _PointEnv:
    """Gym-like toy: a point on a line, reward -|x|, terminal when |x| > 2, time limit of 25 steps."""

    def __init__(self, discrete, seed=0):
        self._rs = np.random.RandomState(seed)
        self.observation_space = _Box((3,), self._rs)
        self.action_space = _Discrete(3, self._rs) if discrete else _Box((1,), self._rs)
        self._discrete = discrete

    def reset(self):
        self._x, self._t = float(self._rs.uniform(-1, 1)), 0
        return np.array([self._x, 0.0, 1.0])

    def step(self, action):
        a = (int(action) - 1) * 0.3 if self._discrete else float(np.asarray(action).reshape(-1)[0]) * 0.3
        self._x += a + 0.05 * float(self._rs.randn())
        self._t += 1
        obs = np.array([self._x, self._t / 25.0, 1.0])
        if abs(self._x) > 2:
            return obs, -abs(self._x), True, {}
        if self._t == 25:
            return obs, -abs(self._x), True, {"TimeLimit.truncated": True}
        return obs, -abs(self._x), False, {}


@pytest.mark.parametrize("name", ["ddpg", "sac", "dqn_qr"])
def test_fit_online_runs_the_update_path_from_the_hbm_buffer(name):
    """`algo.fit_online(env, buffer, explorer, ...)` (algos/base.py:161-247 -> online/iterators.py:99-287): environment
    steps into the HBM buffer, minibatches gathered on the device, one update per `update_interval` steps once the
    buffer holds more than a batch."""
    from d3rlpy_b200.algos import DDPG, DQN, SAC
    from d3rlpy_b200.online import LinearDecayEpsilonGreedy, NormalNoise, ReplayBuffer

    H = [32, 32]
    if name == "ddpg":
        algo, explorer = DDPG(actor_encoder_factory=H, critic_encoder_factory=H, batch_size=32, n_steps=2), NormalNoise()
    elif name == "sac":
        algo, explorer = SAC(actor_encoder_factory=H, critic_encoder_factory=H, batch_size=32), None
    else:
        algo = DQN(encoder_factory=H, q_func_factory="qr", batch_size=32, target_update_interval=50)
        explorer = LinearDecayEpsilonGreedy(1.0, 0.1, 200)
    env = _PointEnv(discrete=name == "dqn_qr")
    buffer = ReplayBuffer(200, env=env)
    np.random.seed(0)
    hist = algo.fit_online(env, buffer, explorer, n_steps=400, n_steps_per_epoch=100, update_interval=2,
                           update_start_step=40, random_steps=20)
    assert len(hist) == 4 and len(buffer) == 200
    # first update at the first even step > 40 with more than 32 transitions stored
    assert algo.grad_step == (400 - 40) // 2
    for h in hist:
        assert "rollout_return" in h and all(np.isfinite(v) for v in h.values())
    keys = {"ddpg": {"critic_loss", "actor_loss"}, "sac": {"critic_loss", "actor_loss", "temp", "temp_loss"},
            "dqn_qr": {"loss"}}[name]
    assert keys <= set(hist[-1])
