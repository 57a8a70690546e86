"""HBM-resident replay buffer + drop-in minibatch types.

Mirrors the public surface of the reference's only native module `d3rlpy.dataset`
(d3rlpy/dataset.pyx; stub d3rlpy/dataset.pyi:44-67) for the part on the hot path:
`MDPDataset -> episodes -> transitions -> TransitionMiniBatch(transitions, n_frames, n_steps,
gamma)` with ndarray properties.  Where the reference walks a `shared_ptr` linked list and
memcpy's per sample on one CPU core, this keeps step-indexed arrays plus 16 B of metadata per
transition in HBM and assembles the minibatch with one coalesced gather kernel (K1/K1b) that is
bit-exact with the reference for observations, frame stacks, actions, terminals and n_steps.
There is no CPU path: building a TransitionMiniBatch needs the CUDA library and a GPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import D3BError, lib


def _transition_meta(terminals: np.ndarray, episode_terminals: np.ndarray):
    """Per-transition {step, episode_start, episode_last_transition_step, flags}
    (flat restatement of _to_transitions, d3rlpy/dataset.pyx:70-116): a terminal episode yields
    one transition per step, a truncated episode drops its last step."""
    ends = np.nonzero(episode_terminals)[0]
    starts = np.concatenate([[0], ends[:-1] + 1]) if len(ends) else np.zeros(0, np.int64)
    meta = []
    ep_ranges = []
    t0 = 0
    for s, e in zip(starts, ends):
        is_term = bool(terminals[e])
        n_tr = (e - s + 1) if is_term else (e - s)
        if n_tr <= 0:
            ep_ranges.append((t0, t0))
            continue
        m = np.empty((n_tr, 4), np.int32)
        m[:, 0] = np.arange(s, s + n_tr)
        m[:, 1] = s
        m[:, 2] = s + n_tr - 1
        m[:, 3] = 0
        if is_term:
            m[-1, 3] = 3  # bit 0: terminal, bit 1: next observation is the zero dummy (csrc/gather.cu)
        meta.append(m)
        ep_ranges.append((t0, t0 + n_tr))
        t0 += n_tr
    meta = np.concatenate(meta, axis=0) if meta else np.zeros((0, 4), np.int32)
    return np.ascontiguousarray(meta), ep_ranges


class DeviceReplay:
    """Step-indexed arrays + transition metadata resident in HBM (DESIGN.md §HBM layout)."""

    def __init__(self, dataset: "MDPDataset", device):
        if not torch.cuda.is_available():
            raise D3BError("DeviceReplay needs a CUDA device; there is no CPU fallback")
        self.dataset = dataset
        self.device = torch.device(device)
        obs = dataset.observations
        self.is_image = obs.ndim == 4
        self.obs = torch.from_numpy(np.ascontiguousarray(obs)).to(self.device)
        self.discrete = dataset.is_action_discrete()
        act = dataset.actions.astype(np.int32 if self.discrete else np.float32)
        self.actions = torch.from_numpy(np.ascontiguousarray(act)).to(self.device)
        self.rewards = torch.from_numpy(np.ascontiguousarray(dataset.rewards, dtype=np.float32)).to(self.device)
        self.meta = torch.from_numpy(dataset._meta).to(self.device)
        self.n_transitions = dataset._meta.shape[0]
        self.obs_shape = tuple(obs.shape[1:])
        self.act_dim = dataset.get_action_size() if self.discrete else act.shape[1]
        self.stream = torch.cuda.Stream(device=self.device)
        self._scaler_cache = {}

    def __len__(self):
        return self.n_transitions

    def scaler_tensors(self, scaler):
        return scaler_device_tensors(self._scaler_cache, scaler, self.device)


def scaler_device_tensors(cache: dict, scaler, device):
    """(subtrahend, divisor, eps) of `(x - s) / (d + eps)` as device tensors, cached per scaler OBJECT (the entry holds
    a reference, so an id cannot be recycled while it is cached).  StandardScaler: (mean, std, eps); MinMaxScaler:
    (min, max - min, 0).  Shared by the offline replay and the online ReplayBuffer."""
    key = id(scaler)
    hit = cache.get(key)
    if hit is None or hit[0] is not scaler:
        sub, div, eps = scaler.affine_f32()
        hit = (scaler, torch.tensor(sub, device=device), torch.tensor(div, device=device), eps)
        cache[key] = hit
    return hit[1], hit[2], hit[3]


class Transition:
    """Handle onto one transition of an MDPDataset (d3rlpy/dataset.pyx:792-1033 surface)."""

    __slots__ = ("_ds", "_t")

    def __init__(self, ds: "MDPDataset", t: int):
        self._ds, self._t = ds, t

    def _m(self):
        return self._ds._meta[self._t]

    def get_observation_shape(self):
        return self._ds.get_observation_shape()

    def get_action_size(self):
        return self._ds.get_action_size()

    @property
    def is_discrete(self):
        return self._ds.is_action_discrete()

    @property
    def observation(self):
        return self._ds.observations[self._m()[0]]

    @property
    def action(self):
        a = self._ds.actions[self._m()[0]]
        return int(a) if self._ds.is_action_discrete() else a

    @property
    def reward(self):
        return float(self._ds.rewards[self._m()[0]])

    @property
    def next_observation(self):
        m = self._m()
        if m[3] & 2:
            return np.zeros_like(self._ds.observations[m[0]])  # dummy after terminal (dataset.pyx:86-90)
        return self._ds.observations[m[0] + 1]

    @property
    def terminal(self):
        return float(self._m()[3] & 1)

    @property
    def prev_transition(self):
        m = self._m()
        return None if m[0] == m[1] else Transition(self._ds, self._t - 1)

    @property
    def next_transition(self):
        m = self._m()
        return None if m[0] == m[2] else Transition(self._ds, self._t + 1)


class Episode:
    """d3rlpy/dataset.pyx:602-783 surface: per-episode array views + transitions."""

    def __init__(self, ds: "MDPDataset", index: int, s: int, e: int, tr_range):
        self._ds, self._index, self._s, self._e, self._tr = ds, index, s, e, tr_range

    @property
    def observations(self):
        return self._ds.observations[self._s:self._e + 1]

    @property
    def actions(self):
        return self._ds.actions[self._s:self._e + 1]

    @property
    def rewards(self):
        return self._ds.rewards[self._s:self._e + 1]

    @property
    def terminal(self):
        return float(self._ds.terminals[self._e])

    @property
    def transitions(self) -> List[Transition]:
        return [Transition(self._ds, t) for t in range(*self._tr)]

    def build_transitions(self) -> None:
        """dataset.pyx:720-734: transitions are handles into the dataset's table, nothing to build."""

    def compute_return(self):
        """Sum of the episode's step rewards, the dropped last step of a timed-out episode included
        (dataset.pyx:763-774)."""
        return np.sum(self.rewards)

    def __len__(self):
        return self._tr[1] - self._tr[0]

    def size(self):
        return len(self)

    def __getitem__(self, index):
        return self.transitions[index]

    def __iter__(self):
        return iter(self.transitions)

    def get_observation_shape(self):
        return self._ds.get_observation_shape()

    def get_action_size(self):
        return self._ds.get_action_size()


class MDPDataset:
    """d3rlpy/dataset.pyx:125-599 constructor and accessors (no HDF5 dump/load — outside the hot path)."""

    def __init__(self, observations, actions, rewards, terminals, episode_terminals=None, discrete_action=None):
        observations = np.asarray(observations)
        assert observations.dtype in (np.uint8, np.float32) or observations.ndim == 2
        if observations.dtype != np.uint8:
            observations = observations.astype(np.float32)
        for a in (observations, actions, rewards, terminals):   # check nan (dataset.pyx:207-210); integers cannot be
            a = np.asarray(a)
            assert a.dtype.kind != "f" or not np.isnan(a).any()
        self._observations = np.ascontiguousarray(observations)
        actions = np.asarray(actions)
        if discrete_action is None:  # _check_discrete_action (dataset.pyx:119-122)
            discrete_action = bool(np.all(np.asarray(actions, np.float32) == np.asarray(actions, np.int32)))
        self._discrete = discrete_action
        self._actions = actions.reshape(-1).astype(np.int32) if discrete_action else actions.astype(np.float32)
        self._rewards = np.asarray(rewards, dtype=np.float32).reshape(-1)
        self._terminals = np.asarray(terminals, dtype=np.float32).reshape(-1)
        self._episode_terminals = self._terminals if episode_terminals is None else \
            np.asarray(episode_terminals, dtype=np.float32).reshape(-1)
        self._rebuild()

    def _rebuild(self) -> None:
        """Transition table of the current step arrays; HBM replicas of the old arrays are dropped."""
        self._meta, self._ep_ranges = _transition_meta(self._terminals, self._episode_terminals)
        self._replays = {}

    def build_episodes(self) -> None:
        """dataset.pyx:575-590: episodes are views computed on access, nothing to build."""

    def append(self, observations, actions, rewards, terminals, episode_terminals=None) -> None:
        """dataset.pyx:424-485: new steps behind the existing ones; the transition table is rebuilt and the replay is
        uploaded again on its next use."""
        import warnings

        observations, actions = np.asarray(observations), np.asarray(actions)
        for observation, action in zip(observations, actions):
            assert observation.shape == self.get_observation_shape(), \
                f"Observation shape must be {self.get_observation_shape()}."
            if self._discrete:
                if int(action) >= self.get_action_size():
                    warnings.warn(f"New action size is higher than {self.get_action_size()}.")
            else:
                assert action.shape == (self.get_action_size(),), f"Action size must be {self.get_action_size()}."
        obs_dtype = self._observations.dtype
        self._observations = np.ascontiguousarray(np.vstack([self._observations, observations]).astype(obs_dtype))
        if self._discrete:
            self._actions = np.hstack([self._actions, actions.reshape(-1)]).astype(np.int32)
        else:
            self._actions = np.vstack([self._actions, actions]).astype(np.float32)
        if episode_terminals is None:
            episode_terminals = terminals
        own_episode_terminals = self._episode_terminals
        self._rewards = np.hstack([self._rewards, np.asarray(rewards).reshape(-1)]).astype(np.float32)
        self._terminals = np.hstack([self._terminals, np.asarray(terminals).reshape(-1)]).astype(np.float32)
        self._episode_terminals = np.hstack([own_episode_terminals,
                                             np.asarray(episode_terminals).reshape(-1)]).astype(np.float32)
        self._rebuild()

    def extend(self, dataset) -> None:
        """dataset.pyx:487-505."""
        assert self.is_action_discrete() == dataset.is_action_discrete(), "Dataset must have discrete action-space."
        assert self.get_observation_shape() == dataset.get_observation_shape(), \
            f"Observation shape must be {self.get_observation_shape()}"
        self.append(dataset.observations, dataset.actions, dataset.rewards, dataset.terminals,
                    dataset.episode_terminals)

    def compute_stats(self):
        """dataset.pyx:336-422: return / reward / action / observation statistics with the reference's keys."""
        episode_returns = [episode.compute_return() for episode in self.episodes]
        stats = {
            "return": {"mean": np.mean(episode_returns), "std": np.std(episode_returns),
                       "min": np.min(episode_returns), "max": np.max(episode_returns),
                       "histogram": np.histogram(episode_returns, bins=20)},
            "reward": {"mean": np.mean(self._rewards), "std": np.std(self._rewards), "min": np.min(self._rewards),
                       "max": np.max(self._rewards), "histogram": np.histogram(self._rewards, bins=20)},
        }
        if not self._discrete:
            stats["action"] = {"mean": np.mean(self.actions, axis=0), "std": np.std(self.actions, axis=0),
                               "min": np.min(self.actions, axis=0), "max": np.max(self.actions, axis=0),
                               "histogram": [np.histogram(self.actions[:, i], bins=20)
                                             for i in range(self.get_action_size())]}
        else:
            freqs = [(self.actions == i).sum() for i in range(self.get_action_size())]
            stats["action"] = {"histogram": [freqs, np.arange(self.get_action_size())]}
        stats["observation"] = {"mean": np.mean(self.observations, axis=0), "std": np.std(self.observations, axis=0),
                                "min": np.min(self.observations, axis=0), "max": np.max(self.observations, axis=0)}
        return stats

    def __getitem__(self, index):
        return self.episodes[index]

    def __iter__(self):
        return iter(self.episodes)

    observations = property(lambda self: self._observations)
    actions = property(lambda self: self._actions)
    rewards = property(lambda self: self._rewards)
    terminals = property(lambda self: self._terminals)
    episode_terminals = property(lambda self: self._episode_terminals)

    def is_action_discrete(self) -> bool:
        return self._discrete

    def get_observation_shape(self):
        return tuple(self._observations.shape[1:])

    def get_action_size(self) -> int:
        return int(self._actions.max()) + 1 if self._discrete else int(self._actions.shape[1])

    def __len__(self):
        return len(self._ep_ranges)

    def size(self):
        return len(self)

    @property
    def episodes(self) -> List[Episode]:
        ends = np.nonzero(self._episode_terminals)[0]
        starts = np.concatenate([[0], ends[:-1] + 1]) if len(ends) else []
        return [Episode(self, i, int(s), int(e), self._ep_ranges[i]) for i, (s, e) in enumerate(zip(starts, ends))]

    def transitions(self) -> List[Transition]:
        return [Transition(self, t) for t in range(self._meta.shape[0])]

    def transition_observations(self) -> np.ndarray:
        return self._observations[self._meta[:, 0]]

    def device_replay(self, device=None) -> DeviceReplay:
        device = torch.device(device if device is not None else "cuda:0")
        key = str(device)
        if key not in self._replays:
            self._replays[key] = DeviceReplay(self, device)
        return self._replays[key]


class TransitionMiniBatch:
    """`TransitionMiniBatch(transitions, n_frames=1, n_steps=1, gamma=0.99)` (dataset.pyx:1139-1217;
    stub dataset.pyi:44-67).  The six arrays are gathered on the GPU into one contiguous HBM buffer
    (`_device_batch`, consumed directly by `algo.update`) and exposed as numpy on demand."""

    def __init__(self, transitions: Sequence[Transition], n_frames: int = 1, n_steps: int = 1, gamma: float = 0.99):
        assert len(transitions) > 0
        ds = transitions[0]._ds
        assert all(t._ds is ds for t in transitions), "transitions must come from one MDPDataset"
        idx = np.fromiter((t._t for t in transitions), dtype=np.int64, count=len(transitions))
        self._transitions = list(transitions)
        self._build(ds.device_replay(), idx, n_frames, n_steps, gamma, None)

    @classmethod
    def from_indices(cls, replay: DeviceReplay, indices, n_frames=1, n_steps=1, gamma=0.99, scaler=None,
                     out=None):
        self = cls.__new__(cls)
        self._transitions = None
        self._build(replay, np.ascontiguousarray(indices, dtype=np.int64), n_frames, n_steps, gamma, scaler, out)
        return self

    def _build(self, replay: DeviceReplay, idx: np.ndarray, n_frames, n_steps, gamma, scaler, out=None):
        from .algos.torch.base import DeviceBatch

        L = lib()
        self._replay, self._indices = replay, idx
        B = idx.shape[0]
        stack = replay.is_image and n_frames > 1
        if replay.is_image:
            c, h, w = replay.obs_shape
            shape = (n_frames * c, h, w) if stack else (c, h, w)
            db = out or DeviceBatch(B, 0, replay.act_dim, replay.device, pixel_shape=shape, discrete=replay.discrete)
        else:
            db = out or DeviceBatch(B, replay.obs_shape[0], replay.act_dim, replay.device, discrete=replay.discrete)
        st = replay.stream.cuda_stream
        idx_host = torch.from_numpy(idx).pin_memory()
        idx_dev = torch.empty(B, dtype=torch.int64, device=replay.device)
        L.copy_h2d(idx_dev.data_ptr(), idx_host.data_ptr(), 8 * B, st)
        sc = (None, None, 0.0)
        self.scaled = set()   # which TorchMiniBatch transforms (torch_utility.py:179-185) the buffer already carries
        if scaler is not None and not replay.is_image and hasattr(scaler, "affine_f32"):
            m, s, e = replay.scaler_tensors(scaler)
            sc = (m.data_ptr(), s.data_ptr(), e)
            self.scaled.add("obs")
        self._act_i32 = None
        if replay.discrete:
            self._act_i32 = torch.empty(B, dtype=torch.int32, device=replay.device)
            act_out = self._act_i32.data_ptr()
        else:
            act_out = db.ptr("act")
        if replay.is_image:
            fb = int(np.prod(replay.obs_shape))
            L.gather_frames(replay.obs.data_ptr(), fb, replay.meta.data_ptr(), idx_dev.data_ptr(), B,
                            n_frames if stack else 1, n_steps, db.ptr("obs"), db.ptr("next_obs"), st)
            L.gather_vector(None, 0, replay.actions.data_ptr(), replay.act_dim, int(replay.discrete),
                            replay.rewards.data_ptr(), replay.meta.data_ptr(), idx_dev.data_ptr(), B, n_steps,
                            float(gamma), None, act_out, db.ptr("rew"), None, db.ptr("term"), db.ptr("nsteps"), None,
                            None, 0.0, st)
        else:
            L.gather_vector(replay.obs.data_ptr(), replay.obs_shape[0], replay.actions.data_ptr(), replay.act_dim,
                            int(replay.discrete), replay.rewards.data_ptr(), replay.meta.data_ptr(),
                            idx_dev.data_ptr(), B, n_steps, float(gamma), db.ptr("obs"), act_out, db.ptr("rew"),
                            db.ptr("next_obs"), db.ptr("term"), db.ptr("nsteps"), sc[0], sc[1], sc[2], st)
        if replay.discrete:
            with torch.cuda.stream(replay.stream):
                db.view("act").copy_(self._act_i32)  # float32 copy for the update path (torch_utility.py:146-149)
        L.stream_sync(st)
        self._keep = (idx_host, idx_dev)
        self._device_batch = db

    # ---- ndarray properties (D2H on demand)
    def _np(self, name):
        return self._device_batch.view(name).detach().cpu().numpy()

    @property
    def observations(self):
        return self._np("obs")

    @property
    def next_observations(self):
        return self._np("next_obs")

    @property
    def actions(self):
        if self._act_i32 is not None:
            return self._act_i32.cpu().numpy()
        return self._np("act")

    @property
    def rewards(self):
        return self._np("rew")

    @property
    def terminals(self):
        return self._np("term")

    @property
    def n_steps(self):
        return self._np("nsteps")

    @property
    def transitions(self):
        if self._transitions is None:
            self._transitions = [Transition(self._replay.dataset, int(t)) for t in self._indices]
        return self._transitions

    def size(self):
        return len(self._indices)

    def __len__(self):
        return len(self._indices)

    def __getitem__(self, i):
        return self.transitions[i]

    def __iter__(self):
        return iter(self.transitions)
