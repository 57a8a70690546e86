"""Recording fakes (environment, algorithm, buffer) to pin the ORDER of calls of an online training loop: the same
objects are driven by the unmodified reference `train_single_env` (tests/golden/make_online_loop_trace.py) and by
d3rlpy_b200.online.train_single_env (tests/test_online_loop_cpu.py)."""
import numpy as np


class _Space:
    def __init__(self, shape, trace):
        self.shape, self._trace = shape, trace

    def sample(self):
        self._trace.append(["action_space.sample"])
        return np.array([0.25], np.float32)


class FakeEnv:
    """Episodes of 7 steps ending in a terminal; every third episode is cut at 5 steps by a time limit instead."""

    def __init__(self, trace):
        self._trace = trace
        self.observation_space = _Space((2,), trace)
        self.action_space = _Space((1,), trace)
        self._episode, self._t = -1, 0

    def reset(self):
        self._episode += 1
        self._t = 0
        self._trace.append(["env.reset", self._episode])
        return np.array([self._t, self._episode], np.float64)

    def step(self, action):
        self._t += 1
        self._trace.append(["env.step", round(float(np.asarray(action).reshape(-1)[0]), 4)])
        obs = np.array([self._t, self._episode], np.float64)
        if self._episode % 3 == 2:
            done = self._t == 5
            return obs, 1.0, done, ({"TimeLimit.truncated": True} if done else {})
        return obs, 0.5, self._t == 7, {}


class FakeBuffer:
    def __init__(self, trace):
        self._trace, self._n = trace, 0

    def __len__(self):
        return self._n

    def append(self, observation, action, reward, terminal, clip_episode=None):
        assert observation.dtype == np.float32
        self._n += 1
        self._trace.append(["buffer.append", [float(v) for v in observation], float(reward), bool(terminal),
                            bool(clip_episode)])

    def sample(self, batch_size, n_frames, n_steps, gamma):
        self._trace.append(["buffer.sample", batch_size, n_frames, n_steps, gamma])
        return "batch"

    def clip_episode(self):
        self._trace.append(["buffer.clip_episode"])


class FakeExplorer:
    def __init__(self, trace):
        self._trace = trace

    def sample(self, algo, x, step):
        self._trace.append(["explorer.sample", list(x.shape), step])
        return np.array([[0.75]], np.float32)


class FakeAlgo:
    batch_size, n_frames, n_steps, gamma = 4, 1, 1, 0.99
    scaler = action_scaler = None
    impl = object()   # already built

    def __init__(self, trace):
        self._trace, self._updates = trace, 0

    def sample_action(self, x):
        self._trace.append(["algo.sample_action", list(np.asarray(x).shape)])
        return np.array([[0.5]], np.float32)

    def update(self, batch):
        self._updates += 1
        self._trace.append(["algo.update", batch])
        return {"loss": float(self._updates)}

    # what only the reference's loop asks for
    def set_active_logger(self, logger):
        pass

    def save_params(self, logger):
        pass

    def save_model(self, fname):
        pass


CONFIGS = {
    "plain": dict(n_steps=40, n_steps_per_epoch=10, update_interval=1, update_start_step=0, random_steps=0),
    "delayed": dict(n_steps=45, n_steps_per_epoch=15, update_interval=3, update_start_step=12, random_steps=6),
    "explorer": dict(n_steps=30, n_steps_per_epoch=10, update_interval=2, update_start_step=0, random_steps=4,
                     explorer=True),
    "no_timelimit": dict(n_steps=30, n_steps_per_epoch=10, update_interval=1, update_start_step=0, random_steps=0,
                         timelimit_aware=False),
}


def run(train_single_env, cfg, **extra):
    trace = []
    cfg = dict(cfg)
    explorer = FakeExplorer(trace) if cfg.pop("explorer", False) else None
    epochs = []
    train_single_env(FakeAlgo(trace), FakeEnv(trace), FakeBuffer(trace), explorer=explorer,
                     callback=lambda algo, epoch, total_step: epochs.append([epoch, total_step]), **cfg, **extra)
    return {"trace": trace, "callback": epochs}
