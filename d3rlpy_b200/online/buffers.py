"""Online replay buffer in HBM: same surface as d3rlpy.online.buffers.ReplayBuffer (d3rlpy/online/buffers.py:18-335)
— `ReplayBuffer(maxlen, env=None, episodes=None)`, `append`, `append_episode`, `clip_episode`, `sample`, `size`,
`__len__` — feeding the same gather kernels (csrc/gather.cu) as the offline sampler.

Where the reference keeps a FIFO of linked `Transition` objects (containers.py:14-80) on the host and assembles every
minibatch with per-sample memcpy, this keeps two append-only device logs:

  steps        O[.] / A[.] / R[.]                 one row per `append` call (or per step of an appended Episode)
  transitions  int32 {step, episode_start, episode_last_transition, flags}   (the layout DeviceReplay uses)

and a few host integers.  The FIFO is a window over the transition log: the `maxlen` newest transitions are the live
ones, physical ring slot `i` of the reference's queue holds transition number `i + maxlen * ((count - 1 - i) // maxlen)`,
so `sample` draws `np.random.choice(len, batch_size)` from the same numpy stream as the reference and gathers exactly
the same transitions.  Dropped transitions stay readable for as long as a later transition of their episode is live
(the reference's drop callback only unlinks whole episodes, buffers.py:30-33): n-step returns and frame stacks see
them.  When a log fills up, the live range (from the first step of the oldest live episode) is moved to the front and
the indices are rebased; `append` itself only writes host staging rows, which reach the device in one copy per
array when `sample` is called.  There is no CPU path.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .._lib import D3BError
from ..dataset import TransitionMiniBatch


class _ReplayView:
    """What TransitionMiniBatch.from_indices needs from a replay (the attributes of dataset.DeviceReplay)."""

    dataset = None


class ReplayBuffer:
    def __init__(self, maxlen: int, env=None, episodes: Optional[List] = None, device="cuda:0",
                 stage_steps: int = 4096):
        if not torch.cuda.is_available():
            raise D3BError("ReplayBuffer needs a CUDA device; there is no CPU fallback")
        if env is not None:
            observation_shape = tuple(env.observation_space.shape)
            space = env.action_space
            discrete = hasattr(space, "n")
            action_size = int(space.n) if discrete else int(space.shape[0])
            obs_dtype = np.uint8 if len(observation_shape) == 3 else np.float32
        elif episodes:
            observation_shape = tuple(episodes[0].get_observation_shape())
            action_size = int(episodes[0].get_action_size())
            discrete = np.asarray(episodes[0].actions).ndim == 1
            obs_dtype = np.asarray(episodes[0].observations).dtype
        else:
            raise ValueError("env or episodes are required to determine shape.")
        self._maxlen, self._observation_shape, self._action_size = int(maxlen), observation_shape, action_size
        self._discrete = discrete
        self._device = torch.device(device)
        self._obs_dtype = np.uint8 if obs_dtype == np.uint8 else np.float32
        # ---- device logs (grown on demand) and the view the gather path reads
        self._cap_s = 3 * self._maxlen + 64
        self._cap_t = 2 * self._maxlen + 64
        v = self._view = _ReplayView()
        v.device, v.is_image, v.obs_shape = self._device, len(observation_shape) == 3, observation_shape
        v.discrete, v.act_dim = discrete, action_size
        v.stream = torch.cuda.Stream(device=self._device)
        v.obs = torch.zeros((self._cap_s,) + observation_shape,
                            dtype=torch.uint8 if self._obs_dtype == np.uint8 else torch.float32, device=self._device)
        v.actions = torch.zeros((self._cap_s,) if discrete else (self._cap_s, action_size),
                                dtype=torch.int32 if discrete else torch.float32, device=self._device)
        v.rewards = torch.zeros(self._cap_s, dtype=torch.float32, device=self._device)
        v.meta = torch.zeros(self._cap_t, 4, dtype=torch.int32, device=self._device)
        v.n_transitions = 0
        v._scaler_cache = {}
        v.scaler_tensors = self._scaler_tensors
        self._meta = np.zeros((self._cap_t, 4), np.int32)   # host mirror of the transition log
        # ---- host staging of the steps appended since the last flush
        self._stage_cap = int(stage_steps)
        self._st_obs = np.zeros((self._stage_cap,) + observation_shape, self._obs_dtype)
        self._st_act = np.zeros((self._stage_cap,) if discrete else (self._stage_cap, action_size),
                                np.int32 if discrete else np.float32)
        self._st_rew = np.zeros(self._stage_cap, np.float32)
        self._n_staged = 0
        # ---- counters.  Log positions are relative to the last compaction.
        self._n_steps = 0          # steps in the log (device rows + staged rows)
        self._n_trans = 0          # transitions in the log
        self._count = 0            # transitions appended since construction (FIFO numbering)
        self._t_base = 0           # FIFO number of log position 0
        self._dirty_lo = 0         # first transition-log row the device copy does not have yet
        self._ep_start: Optional[int] = None   # first step of the episode being recorded
        self._ep_first_t = 0       # first transition-log row of that episode
        if episodes:
            for episode in episodes:
                self.append_episode(episode)

    # ------------------------------------------------------------------ reference surface
    def __len__(self) -> int:
        return min(self._count, self._maxlen)

    def size(self) -> int:
        return len(self)

    def append(self, observation, action, reward: float, terminal: float, clip_episode: Optional[bool] = None) -> None:
        """buffers.py:254-308: the step appended BEFORE this one becomes a transition whose `terminal` is this call's;
        a terminal step adds the closing transition with the zero next observation; `clip_episode` ends the episode."""
        if clip_episode is None:
            clip_episode = bool(terminal)
        observation = np.asarray(observation)
        assert observation.shape == self._observation_shape
        if isinstance(action, np.ndarray) and action.ndim > 0:
            assert action.shape[0] == self._action_size
        else:
            action = int(action)
            assert action < self._action_size
        assert not (terminal and not clip_episode)  # not allow terminal=True and clip_episode=False
        s = self._push_step(observation, action, reward)
        if self._ep_start is None:
            self._ep_start, self._ep_first_t = s, self._n_trans
        else:
            self._push_transition(s - 1 - self._ep_start, 1 if terminal else 0)
        if clip_episode:
            if terminal:
                # _add_last_step (buffers.py:317-335): terminal + zero dummy
                self._push_transition(self._n_steps - 1 - self._ep_start, 3)
            self.clip_episode()

    def append_episode(self, episode) -> None:
        """buffers.py:55-66: every transition of an Episode (`_to_transitions`, dataset.pyx:70-116)."""
        assert tuple(episode.get_observation_shape()) == self._observation_shape
        assert episode.get_action_size() == self._action_size
        assert self._ep_start is None, "append_episode in the middle of a recorded episode"
        obs, act, rew = np.asarray(episode.observations), np.asarray(episode.actions), np.asarray(episode.rewards)
        n, terminal = len(rew), bool(episode.terminal)
        for i in range(n):
            s = self._push_step(obs[i], act[i] if not self._discrete else int(act[i]), rew[i])
            if i == 0:  # set before the other steps arrive: a compaction in between rebases it with the logs
                self._ep_start, self._ep_first_t = s, self._n_trans
        for i in range(n if terminal else n - 1):
            self._push_transition(i, 3 if (terminal and i == n - 1) else 0)
        self.clip_episode()

    def clip_episode(self) -> None:
        self._ep_start = None

    def sample(self, batch_size: int, n_frames: int = 1, n_steps: int = 1, gamma: float = 0.99) -> TransitionMiniBatch:
        """BasicSampleMixin.sample (buffers.py:211-221): slots drawn from the global numpy stream, gathered on the GPU."""
        slots = np.random.choice(len(self), batch_size)
        return self.sample_slots(slots, n_frames, n_steps, gamma)

    def sample_slots(self, slots, n_frames: int = 1, n_steps: int = 1, gamma: float = 0.99, scaler=None,
                     out=None) -> TransitionMiniBatch:
        slots = np.asarray(slots, np.int64).reshape(-1)
        assert slots.size > 0 and slots.min() >= 0 and slots.max() < len(self)
        number = slots + self._maxlen * ((self._count - 1 - slots) // self._maxlen)   # FIFOQueue ring slot -> append number
        self.flush()
        return TransitionMiniBatch.from_indices(self._view, number - self._t_base, n_frames, n_steps, gamma, scaler, out)

    @property
    def observation_shape(self):
        return self._observation_shape

    @property
    def action_size(self) -> int:
        return self._action_size

    def to_mdp_dataset(self):
        raise NotImplementedError("to_mdp_dataset is outside the update path (the logs live in HBM)")

    # ------------------------------------------------------------------ logs
    def _push_step(self, observation, action, reward) -> int:
        if self._n_staged == self._stage_cap:
            self.flush()
        if self._n_steps == self._cap_s:
            self._make_room()
        i = self._n_staged
        self._st_obs[i] = observation
        self._st_act[i] = action
        self._st_rew[i] = reward
        self._n_staged += 1
        self._n_steps += 1
        return self._n_steps - 1

    def _push_transition(self, offset: int, flags: int) -> None:
        """Transition of step `episode start + offset` of the episode being recorded."""
        if self._n_trans == self._cap_t:
            self._make_room()
        step = self._ep_start + offset   # after any compaction: log positions are rebased, offsets are not
        t = self._n_trans
        self._meta[t] = (step, self._ep_start, step, flags)
        self._meta[self._ep_first_t:t, 2] = step   # the episode's last transition moved on (k = min(n_steps, last - g + 1))
        self._dirty_lo = min(self._dirty_lo, self._ep_first_t)
        self._n_trans += 1
        self._count += 1

    def flush(self) -> None:
        """Staged steps and the transition rows that changed since the last flush -> device (one copy per array)."""
        v = self._view
        with torch.cuda.stream(v.stream):
            n = self._n_staged
            if n:
                lo = self._n_steps - n
                v.obs[lo:lo + n].copy_(torch.from_numpy(self._st_obs[:n]))
                v.actions[lo:lo + n].copy_(torch.from_numpy(self._st_act[:n]))
                v.rewards[lo:lo + n].copy_(torch.from_numpy(self._st_rew[:n]))
                self._n_staged = 0
            if self._dirty_lo < self._n_trans:
                v.meta[self._dirty_lo:self._n_trans].copy_(torch.from_numpy(self._meta[self._dirty_lo:self._n_trans]))
                self._dirty_lo = self._n_trans
        v.stream.synchronize()
        v.n_transitions = self._n_trans

    def _make_room(self) -> None:
        """Moves the live range of both logs to the front (rebasing step indices); doubles a log that is still full."""
        self.flush()
        v = self._view
        live = min(self._count, self._maxlen)
        lo_t = self._n_trans - live                       # oldest live transition
        if self._ep_start is not None:
            lo_t = min(lo_t, self._ep_first_t)            # an episode longer than maxlen keeps all its rows
        lo_s = int(self._meta[lo_t, 1]) if lo_t < self._n_trans else (self._ep_start if self._ep_start is not None
                                                                      else self._n_steps)
        if self._ep_start is not None:
            lo_s = min(lo_s, self._ep_start)
        with torch.cuda.stream(v.stream):
            if lo_s > 0:
                n = self._n_steps - lo_s
                for name in ("obs", "actions", "rewards"):
                    t = getattr(v, name)
                    t[:n].copy_(t[lo_s:lo_s + n].clone())
                self._n_steps = n
                if self._ep_start is not None:
                    self._ep_start -= lo_s
            if lo_t > 0 or lo_s > 0:
                n = self._n_trans - lo_t
                self._meta[:n] = self._meta[lo_t:lo_t + n]
                self._meta[:n, :3] -= lo_s
                self._n_trans = n
                self._t_base += lo_t
                self._ep_first_t -= lo_t
                self._dirty_lo = 0
            if self._n_steps > self._cap_s // 2:           # still more than half full: grow
                self._cap_s *= 2
                for name in ("obs", "actions", "rewards"):
                    t = getattr(v, name)
                    g = torch.zeros((self._cap_s,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                    g[:self._n_steps].copy_(t[:self._n_steps])
                    setattr(v, name, g)
            if self._n_trans > self._cap_t // 2:
                self._cap_t *= 2
                m = np.zeros((self._cap_t, 4), np.int32)
                m[:self._n_trans] = self._meta[:self._n_trans]
                self._meta = m
                v.meta = torch.zeros(self._cap_t, 4, dtype=torch.int32, device=self._device)
                self._dirty_lo = 0
        v.stream.synchronize()
        self.flush()

    def _scaler_tensors(self, scaler):
        from ..dataset import scaler_device_tensors

        return scaler_device_tensors(self._view._scaler_cache, scaler, self._device)
