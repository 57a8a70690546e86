"""Replays the scripted append sequences of tests/golden/online.npz (tests/golden/make_golden_online.py) into any
buffer with the reference's ReplayBuffer surface."""
from types import SimpleNamespace

import numpy as np

FIELDS = ("observations", "actions", "rewards", "next_observations", "terminals", "n_steps")


def episode_ns(obs, act, rew, terminal, action_size):
    """Duck-typed Episode (observations / actions / rewards / terminal + shape getters)."""
    return SimpleNamespace(observations=obs, actions=act, rewards=rew, terminal=float(terminal),
                           get_observation_shape=lambda: tuple(obs.shape[1:]), get_action_size=lambda: action_size)


def initial_episodes(z, name):
    maxlen, with_init, discrete, asize = [int(v) for v in z[f"{name}/cfg"]]
    obs = z[f"{name}/script/observations"]
    if with_init:
        return [episode_ns(z[f"{name}/init{e}/observations"], z[f"{name}/init{e}/actions"],
                           z[f"{name}/init{e}/rewards"], float(z[f"{name}/init{e}/terminal"]), asize) for e in (0, 1)]
    # the recording gave the reference its shapes through a truncated all-zero 2-step episode (one transition)
    zo = np.zeros((2,) + obs.shape[1:], obs.dtype)
    za = np.zeros(2, np.int32) if discrete else np.zeros((2, asize), np.float32)
    return [episode_ns(zo, za, np.zeros(2, np.float32), 0.0, asize)]


def replay(z, name, buf, sample):
    """Feeds the script; at every recorded check calls sample(buf, B, n_frames, n_steps) under the recorded numpy
    seed and yields (check index, result, reference arrays)."""
    obs, act = z[f"{name}/script/observations"], z[f"{name}/script/actions"]
    rew, term, clip = z[f"{name}/script/rewards"], z[f"{name}/script/terminals"], z[f"{name}/script/clips"]
    checks = z[f"{name}/checks"]
    j = 0
    for i in range(len(rew)):
        buf.append(obs[i], act[i], float(rew[i]), float(term[i]), clip_episode=bool(clip[i]))
        while j < len(checks) and checks[j][0] == i:
            _, seed, B, n_frames, n_steps, size = [int(v) for v in checks[j]]
            assert len(buf) == size, (name, i, len(buf), size)
            np.random.seed(seed)
            got = sample(buf, B, n_frames, n_steps)
            yield j, got, {k: z[f"{name}/ref{j}/{k}"] for k in FIELDS}
            j += 1
    assert j == len(checks)
