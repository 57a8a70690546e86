"""Golden fixtures for the online replay buffer: scripted `append` / `append_episode`
sequences driven through the LIVE unmodified reference `d3rlpy.online.buffers.ReplayBuffer` (ring wrap-around, true
terminals, time-out clips, sampling in the middle of an episode), with the minibatches it samples under fixed numpy
seeds.  The oracle restatement (oracle/sampler.py:OnlineReplay) is checked against the reference while recording.

    python tests/golden/make_golden_online.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference through oracle/ref_import)

from oracle import sampler as osampler  # noqa: E402

FIELDS = ("observations", "actions", "rewards", "next_observations", "terminals", "n_steps")


def script(rs, kind, n_events):
    """Random walk of episodes: lengths 2..14, each ending in a terminal (p = 0.5) or a time-out clip."""
    obs, act, rew, term, clip = [], [], [], [], []
    while len(rew) < n_events:
        n = int(rs.randint(2, 15))
        terminal = rs.rand() < 0.5
        for i in range(n):
            if kind == "image":
                obs.append(rs.randint(0, 256, size=(1, 8, 8)).astype(np.uint8))
                act.append(int(rs.randint(4)))
            elif kind == "vector_disc":
                obs.append(rs.randn(5).astype(np.float32))
                act.append(int(rs.randint(3)))
            else:
                obs.append(rs.randn(5).astype(np.float32))
                act.append(rs.uniform(-1, 1, 2).astype(np.float32))
            rew.append(np.float32(rs.randn()))
            last = i == n - 1
            term.append(1.0 if (last and terminal) else 0.0)
            clip.append(1 if last else 0)
    return obs[:n_events], act[:n_events], rew[:n_events], term[:n_events], clip[:n_events]


def main():
    from d3rlpy.dataset import Episode
    from d3rlpy.online.buffers import ReplayBuffer

    rs = np.random.RandomState(31)
    out, cases = {}, []
    for kind in ("vector_cont", "vector_disc", "image"):
        for with_init in (False, True):
            name = f"{kind}_{'init' if with_init else 'empty'}"
            maxlen, n_events = 40, 150
            discrete = kind != "vector_cont"
            oshape = (1, 8, 8) if kind == "image" else (5,)
            asize = 4 if kind == "image" else 3 if kind == "vector_disc" else 2
            obs, act, rew, term, clip = script(rs, kind, n_events)
            episodes = None
            orc = osampler.OnlineReplay(maxlen, oshape, discrete)
            if with_init:  # one terminal and one truncated offline episode in the buffer from the start
                episodes = []
                for e, terminal in enumerate((1.0, 0.0)):
                    eo, ea, er, _, _ = script(rs, kind, 7)
                    eo, er = np.stack(eo), np.asarray(er, np.float32)
                    ea = np.asarray(ea, np.int32) if discrete else np.stack(ea)
                    episodes.append(Episode(oshape, asize, eo, ea, er, terminal))
                    out[f"{name}/init{e}/observations"], out[f"{name}/init{e}/actions"] = eo, ea
                    out[f"{name}/init{e}/rewards"], out[f"{name}/init{e}/terminal"] = er, np.float32(terminal)
                    orc.append_episode(eo, ea, er, bool(terminal))
                buf = ReplayBuffer(maxlen, episodes=episodes)
            else:
                class _Env:  # shape information only (buffers.py:37-40)
                    pass
                buf = ReplayBuffer(maxlen, episodes=[Episode(oshape, asize, np.zeros((2,) + oshape, obs[0].dtype),
                                                             np.zeros(2, np.int32) if discrete else np.zeros((2, asize), np.float32),
                                                             np.zeros(2, np.float32), 0.0)])
                # the truncated 2-step episode above leaves ONE transition in the buffer; mirror it
                z = np.zeros((2,) + oshape, obs[0].dtype)
                orc.append_episode(z, np.zeros(2, np.int32) if discrete else np.zeros((2, asize), np.float32),
                                   np.zeros(2, np.float32), False)
            checks, j = [], 0
            for i in range(n_events):
                a = act[i] if discrete else act[i]
                buf.append(obs[i], a, float(rew[i]), term[i], clip_episode=bool(clip[i]))
                orc.append(obs[i], a, float(rew[i]), term[i], clip_episode=bool(clip[i]))
                assert len(buf) == len(orc), (name, i, len(buf), len(orc))
                if i in (5, 23, 61, 100, 149):
                    for n_frames, n_steps in ((1, 1), (1, 3), (4, 1), (4, 3)):
                        if kind != "image" and n_frames > 1:
                            continue
                        seed, B = 1000 + j, 32
                        np.random.seed(seed)
                        rb = mg.batch_arrays(buf.sample(B, n_frames, n_steps, 0.99))
                        np.random.seed(seed)
                        ob = orc.sample(B, n_frames, n_steps, 0.99)
                        for k in FIELDS:
                            if k == "rewards":
                                assert np.allclose(rb[k], ob[k], rtol=1e-6, atol=1e-7), (name, i, k)
                            else:
                                assert rb[k].dtype == ob[k].dtype and np.array_equal(rb[k], ob[k]), (name, i, k, n_frames, n_steps)
                            out[f"{name}/ref{j}/{k}"] = rb[k]
                        checks.append((i, seed, B, n_frames, n_steps, len(buf)))
                        j += 1
            out[f"{name}/script/observations"] = np.stack(obs)
            out[f"{name}/script/actions"] = np.asarray(act, np.int32) if discrete else np.stack(act)
            out[f"{name}/script/rewards"] = np.asarray(rew, np.float32)
            out[f"{name}/script/terminals"] = np.asarray(term, np.float32)
            out[f"{name}/script/clips"] = np.asarray(clip, np.int32)
            out[f"{name}/checks"] = np.asarray(checks, np.int64)
            out[f"{name}/cfg"] = np.asarray([maxlen, int(with_init), int(discrete), asize], np.int64)
            cases.append(name)
    out["cases"] = np.array(cases)
    path = os.path.join(HERE, "online.npz")
    np.savez_compressed(path, **out)
    print("online.npz:", cases, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
