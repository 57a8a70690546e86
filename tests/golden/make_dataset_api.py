"""Golden fixture for the MDPDataset / Episode surface around the sampler (append, extend, compute_stats, episode
returns, iteration), recorded from the LIVE unmodified reference's compiled `d3rlpy.dataset` (oracle/_ref).

    python tests/golden/make_dataset_api.py        (build container only: needs /root/reference)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

d3 = mg.d3


def arrays(rs, n, ep, obs=4, act=2, discrete=False, timeout_every=3):
    o = rs.randn(n, obs).astype(np.float32)
    a = rs.randint(0, 3, size=n).astype(np.int32) if discrete else rs.uniform(-1, 1, (n, act)).astype(np.float32)
    r = rs.randn(n).astype(np.float32)
    ept = np.zeros(n, np.float32)
    ept[ep - 1::ep] = 1.0
    ept[-1] = 1.0
    t = ept.copy()
    t[ep - 1::timeout_every * ep] = 0.0
    return o, a, r, t, ept


def describe(ds):
    """Everything a caller can observe, as plain lists."""
    stats = ds.compute_stats()
    out = {"n_episodes": len(ds), "size": ds.size(), "action_size": int(ds.get_action_size()),
           "observation_shape": list(ds.get_observation_shape()), "discrete": bool(ds.is_action_discrete()),
           "episode_sizes": [e.size() for e in ds.episodes], "episode_lens": [len(e) for e in ds],
           "episode_returns": [float(e.compute_return()) for e in ds.episodes],
           "episode_terminal": [float(e.terminal) for e in ds.episodes],
           "n_steps_per_episode": [int(e.observations.shape[0]) for e in ds.episodes],
           "stats": {}}
    for grp, d in stats.items():
        out["stats"][grp] = {}
        for k, v in d.items():
            if k == "histogram":
                if grp == "action" and not ds.is_action_discrete():
                    v = [[np.asarray(h[0]).tolist(), np.asarray(h[1]).tolist()] for h in v]
                else:
                    v = [np.asarray(v[0]).tolist(), np.asarray(v[1]).tolist()]
            else:
                v = np.asarray(v, np.float64).tolist()
            out["stats"][grp][k] = v
    tr = ds.episodes[1].transitions
    out["episode1_first_last"] = [np.asarray(tr[0].observation).tolist(), np.asarray(tr[-1].next_observation).tolist(),
                                  float(tr[-1].terminal), float(ds.episodes[1][0].reward)]
    return out


def main():
    rs = np.random.RandomState(41)
    doc = {}
    for name, discrete in (("continuous", False), ("discrete", True)):
        base = arrays(rs, 60, 12, discrete=discrete)
        more = arrays(rs, 35, 9, discrete=discrete, timeout_every=2)
        other = arrays(rs, 27, 9, discrete=discrete)
        ds = d3.dataset.MDPDataset(*base, discrete_action=discrete)
        case = {"base": describe(ds)}
        ds.append(*more)
        case["appended"] = describe(ds)
        ds.extend(d3.dataset.MDPDataset(*other, discrete_action=discrete))
        case["extended"] = describe(ds)
        case["inputs"] = {k: [np.asarray(x).tolist() for x in v] for k, v in
                          (("base", base), ("more", more), ("other", other))}
        doc[name] = case
    with open(os.path.join(HERE, "dataset_api.json"), "w") as f:
        json.dump(doc, f)
    print("dataset_api.json", os.path.getsize(os.path.join(HERE, "dataset_api.json")) // 1024, "KiB")


if __name__ == "__main__":
    main()
