"""Shared host machinery of the B200 impls: device batch staging, noise arena, metric slots,
step counters and whole-update CUDA-graph capture (K12).

Mirrors the role of TorchImplBase + @torch_api/@train_api (d3rlpy/algos/torch/base.py:20-142,
d3rlpy/torch_utility.py:152-299): turns a TransitionMiniBatch into float32 device tensors
(applying the scalers), and returns losses as numpy scalars.  Unlike the reference it converts the
batch ONCE per update, and reads every metric back with ONE pinned D2H copy.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from ..._lib import D3BError, lib


def _align4(n: int) -> int:
    return (n + 3) // 4 * 4


class DeviceBatch:
    """One contiguous float32 HBM buffer holding the six minibatch arrays
    [obs | next_obs | actions | rewards | terminals | n_steps] (vector observations), mirrored by a
    pinned host staging buffer of the same layout (single H2D copy per update)."""

    def __init__(self, batch: int, obs_dim: int, act_dim: int, device, pixel_shape=None, discrete=False):
        self.B, self.O, self.A = batch, obs_dim, act_dim
        self.pixel_shape = tuple(pixel_shape) if pixel_shape else None
        self.discrete = discrete
        o = 0
        self.off = {}
        fields = [("obs", 0 if pixel_shape else batch * obs_dim), ("next_obs", 0 if pixel_shape else batch * obs_dim),
                  ("act", batch * (1 if discrete else act_dim)), ("rew", batch), ("term", batch), ("nsteps", batch)]
        for name, n in fields:
            self.off[name] = o
            o = _align4(o + n)
        self.nfloat = o
        self.dev = torch.zeros(o, dtype=torch.float32, device=device)
        self.host = torch.zeros(o, dtype=torch.float32).pin_memory() if device.type == "cuda" else torch.zeros(o)
        self.host_np = self.host.numpy()
        if pixel_shape:
            n = batch * int(np.prod(pixel_shape))
            self.pix_dev = torch.zeros(2 * n, dtype=torch.uint8, device=device)
            self.pix_host = torch.zeros(2 * n, dtype=torch.uint8)
            if device.type == "cuda":
                self.pix_host = self.pix_host.pin_memory()
            self.pix_host_np = self.pix_host.numpy()
            self.npix = n
        # shaped views of the staging buffer, one per minibatch property: `stage_host` is six np.copyto calls
        h, B = self.host_np, batch
        shapes = {"act": (B,) if discrete else (B, act_dim), "rew": (B, 1), "term": (B, 1), "nsteps": (B, 1)}
        if pixel_shape:
            self._stage = [("observations", self.pix_host_np[:n].reshape(B, *self.pixel_shape)),
                           ("next_observations", self.pix_host_np[n:].reshape(B, *self.pixel_shape))]
        else:
            shapes.update(obs=(B, obs_dim), next_obs=(B, obs_dim))
            self._stage = [("observations", self._host_view(h, "obs", shapes)),
                           ("next_observations", self._host_view(h, "next_obs", shapes))]
        self._stage += [("actions", self._host_view(h, "act", shapes)), ("rewards", self._host_view(h, "rew", shapes)),
                        ("terminals", self._host_view(h, "term", shapes)), ("n_steps", self._host_view(h, "nsteps", shapes))]

    def _host_view(self, h, name, shapes):
        shape = shapes[name]
        return h[self.off[name]:self.off[name] + int(np.prod(shape))].reshape(shape)

    def ptr(self, name: str) -> int:
        if self.pixel_shape and name in ("obs", "next_obs"):
            return self.pix_dev.data_ptr() + (self.npix if name == "next_obs" else 0)
        return self.dev.data_ptr() + 4 * self.off[name]

    def view(self, name: str) -> torch.Tensor:
        B = self.B
        if self.pixel_shape and name in ("obs", "next_obs"):
            s = self.npix if name == "next_obs" else 0
            return self.pix_dev[s:s + self.npix].view(B, *self.pixel_shape)
        shape = {"obs": (B, self.O), "next_obs": (B, self.O), "act": (B,) if self.discrete else (B, self.A),
                 "rew": (B, 1), "term": (B, 1), "nsteps": (B, 1)}[name]
        n = int(np.prod(shape))
        return self.dev[self.off[name]:self.off[name] + n].view(shape)

    @property
    def h2d_bytes(self) -> int:
        return 4 * self.nfloat + (2 * self.npix if self.pixel_shape else 0)

    def stage_host(self, batch) -> None:
        """numpy TransitionMiniBatch-like -> pinned staging (float32 casts as _convert_to_torch,
        d3rlpy/torch_utility.py:146-149; uint8 frames stay uint8 on the wire)."""
        for name, view in self._stage:
            src = getattr(batch, name)
            if getattr(src, "shape", None) != view.shape:   # (B,) scalars, lists, flattened frames
                src = np.asarray(src).reshape(view.shape)
            np.copyto(view, src, casting="unsafe")


class ImplBase:
    """Common state: device, stream, counters, metric slots, noise arena, graphs."""

    METRICS: Sequence[str] = ()
    N_COUNTERS = 8

    def __init__(self, observation_shape, action_size, use_gpu=0, scaler=None, action_scaler=None,
                 reward_scaler=None, world_size: int = 1, rank: int = 0):
        if not torch.cuda.is_available():
            raise D3BError("d3rlpy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._lib = lib()
        dev_id = 0 if (use_gpu is None or isinstance(use_gpu, bool)) else int(getattr(use_gpu, "get_id", lambda: use_gpu)())
        self._device = torch.device("cuda", dev_id)
        torch.cuda.set_device(self._device)
        self._observation_shape = tuple(observation_shape)
        self._action_size = int(action_size)
        self._scaler, self._action_scaler, self._reward_scaler = scaler, action_scaler, reward_scaler
        self._stream_obj = torch.cuda.Stream(device=self._device)
        self._stream = self._stream_obj.cuda_stream
        self._counters = torch.zeros(self.N_COUNTERS, dtype=torch.int32, device=self._device)
        self._slots = torch.zeros(64, dtype=torch.float32, device=self._device)  # [0:32) metrics, [32:64) sums
        self._slots_host = torch.zeros(64, dtype=torch.float32).pin_memory()
        self._slots_host_np = self._slots_host.numpy()   # the same pinned words, as numpy
        self._graphs: Dict[tuple, int] = {}
        self._graph_nodes: Dict[tuple, int] = {}
        self._batch: Optional[DeviceBatch] = None          # the active minibatch buffer
        self._batches: Dict[int, DeviceBatch] = {}         # impl-owned buffers, one per batch size (stable pointers)
        self._noise_by_b: Dict[int, torch.Tensor] = {}
        self._noise_injected = False
        self._seed = 0
        self._ws: Dict[str, torch.Tensor] = {}
        self.world_size, self.rank = world_size, rank
        if world_size > 1:
            from ... import parallel

            parallel.init(world_size, rank)  # library-owned NCCL communicator (K11)
        self.use_graph = True

    # ------------------------------------------------------------------ small helpers
    @property
    def device(self) -> str:
        return str(self._device)

    @property
    def observation_shape(self):
        return self._observation_shape

    @property
    def action_size(self) -> int:
        return self._action_size

    @property
    def scaler(self):
        return self._scaler

    @property
    def action_scaler(self):
        return self._action_scaler

    @property
    def reward_scaler(self):
        return self._reward_scaler

    def ws(self, name: str, *shape, dtype=torch.float32) -> torch.Tensor:
        """Named workspace.  Keyed by (name, shape, dtype): buffers of different minibatch sizes coexist, so a
        captured graph of one size keeps valid pointers while another size (e.g. a 1-row evaluation) runs."""
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.zeros(*shape, dtype=dtype, device=self._device)
            self._ws[key] = t
            # the zero-fill ran on torch's current stream; our kernels run on the private stream
            torch.cuda.current_stream(self._device).synchronize()
        return t

    def counter_ptr(self, i: int) -> int:
        return self._counters.data_ptr() + 4 * i

    def metric_ptr(self, i: int) -> int:
        return self._slots.data_ptr() + 4 * i

    def sums_ptr(self, i: int) -> int:
        return self._slots.data_ptr() + 4 * (32 + i)

    def zero_slots(self):
        self._lib.memset_zero(self._slots.data_ptr(), 4 * 64, self._stream)

    def sync(self):
        self._lib.stream_sync(self._stream)

    def read_slots_after_program(self) -> np.ndarray:
        """Metrics of the update enqueued by `run_program` (its last node is the pinned D2H of the slots)."""
        if not getattr(self, "_metrics_on_host", False):
            return self.read_slots()  # device-resident batch: the graph has no read-back node
        self.sync()
        return self._slots_host_np

    # Staged copies up to this many bytes run as a kernel over the pinned (device-addressable) host buffer instead of
    # a copy-engine transfer: in front of / behind a 160 us update the node's start latency is what counts
    # (csrc/util.cu: copy_mapped).  D3B_ZEROCOPY_MAX=0 restores cudaMemcpyAsync everywhere.
    zero_copy_max = int(os.environ.get("D3B_ZEROCOPY_MAX", str(256 << 10)))

    def _copy_h2d(self, dst: int, src_pinned: int, nbytes: int) -> None:
        if nbytes <= self.zero_copy_max:
            self._lib.copy_mapped(dst, src_pinned, nbytes, self._stream)
        else:
            self._lib.copy_h2d(dst, src_pinned, nbytes, self._stream)

    def _copy_d2h(self, dst_pinned: int, src: int, nbytes: int) -> None:
        if nbytes <= self.zero_copy_max:
            self._lib.copy_mapped(dst_pinned, src, nbytes, self._stream)
        else:
            self._lib.copy_d2h(dst_pinned, src, nbytes, self._stream)

    def read_slots(self) -> np.ndarray:
        self._copy_d2h(self._slots_host.data_ptr(), self._slots.data_ptr(), 4 * 64)
        self.sync()
        return self._slots_host_np

    # ------------------------------------------------------------------ checkpoints
    def _checkpoint_views(self):
        raise NotImplementedError

    # ------------------------------------------------------------------ deployment export (algos/torch/base.py:86-126)
    POLICY_KIND: str = ""

    def save_policy(self, fname: str) -> None:
        """Greedy policy (observation scaler included) as TorchScript `.pt` / ONNX `.onnx`; see d3rlpy_b200/export.py."""
        from ...export import GreedyPolicy, save_policy

        self.sync()
        pol = getattr(self, "policy", None)
        imit = getattr(self, "imitator", None)
        module = GreedyPolicy(self.POLICY_KIND, policy=pol.state_dict() if pol is not None else None,
                              q=self.q_function.state_dict(), imitator=imit.state_dict() if imit is not None else None,
                              scaler=self._scaler, n_action_samples=getattr(self, "_n_action_samples", 100),
                              action_flexibility=getattr(self, "_action_flexibility", 0.05),
                              n_quantiles=getattr(self, "_n_quantiles", 0), action_scaler=self._action_scaler)
        save_policy(module, self.observation_shape, fname)

    def save_model(self, fname: str) -> None:
        """torch.save({attr: state_dict}) with the reference's attribute names and state_dict keys
        (algos/torch/base.py:137-139, torch_utility.py:97-103), so either side can load the other's file."""
        self.sync()
        torch.cuda.synchronize(self._device)

        def cpu(o):
            if isinstance(o, torch.Tensor):
                return o.detach().cpu().clone()
            if isinstance(o, dict):
                return type(o)((k, cpu(v)) for k, v in o.items())
            if isinstance(o, list):
                return [cpu(v) for v in o]
            return o

        torch.save({k: cpu(v.state_dict()) for k, v in self._checkpoint_views().items()}, fname)

    def load_model(self, fname: str) -> None:
        """set_state_dict (torch_utility.py:106-110): every module / optimizer attribute is restored."""
        chkpt = torch.load(fname, map_location="cpu", weights_only=False)
        for k, v in self._checkpoint_views().items():
            v.load_state_dict(chkpt[k])
        torch.cuda.synchronize(self._device)

    # ------------------------------------------------------------------ batches
    def _make_batch(self, B: int) -> DeviceBatch:
        pixel = self._observation_shape if len(self._observation_shape) == 3 else None
        O = 0 if pixel else self._observation_shape[0]
        return DeviceBatch(B, O, self._action_size, self._device, pixel_shape=pixel, discrete=self.DISCRETE)

    DISCRETE = False

    def device_batch(self, B: int) -> DeviceBatch:
        """The impl-owned minibatch buffer of size B (created once: graphs captured over it stay valid when another
        batch size is used in between, e.g. the 1-row evaluation batches of the online loop)."""
        db = self._batches.get(B)
        if db is None:
            db = self._batches[B] = self._make_batch(B)
            torch.cuda.current_stream(self._device).synchronize()
        self._batch = db
        return db

    def load_batch(self, batch, defer: bool = False) -> DeviceBatch:
        """Accepts a host minibatch (numpy properties, e.g. the reference's TransitionMiniBatch or ours)
        or our device-resident TransitionMiniBatch (already gathered in HBM).  defer=True (whole-update
        graphs): the arrays are only staged in pinned memory; the H2D copy and the scaler become the first nodes
        of the update graph (`run_program`)."""
        self._pending_upload = None
        dev = getattr(batch, "_device_batch", None)
        if dev is not None:
            dev = self._adopt_device_batch(batch, dev)
            self._batch = dev
            return dev
        B = len(batch.rewards) if hasattr(batch, "rewards") else len(batch)
        db = self.device_batch(B)
        db.stage_host(batch)
        if defer:
            self._pending_upload = db
        else:
            self._upload(db)
        return db

    def _adopt_device_batch(self, batch, dev: DeviceBatch) -> DeviceBatch:
        """A minibatch gathered in HBM carries in `batch.scaled` which of the TorchMiniBatch transforms
        (torch_utility.py:179-185) were already applied to it (fused into the gather, or by `fit`).  Whatever is
        missing is applied here.  A buffer owned by the caller (a fresh `ReplayBuffer.sample()` /
        `TransitionMiniBatch(transitions)`, whose numpy properties must keep showing the raw data) is copied device to
        device into the impl-owned buffer of that size first: the update graph is captured once over the impl's
        buffers and replayed for every such batch instead of being re-captured per sample."""
        have = getattr(batch, "scaled", True)   # a bare holder of device buffers: its owner manages the transforms
        have = {"obs", "act_rew"} if have is True else set(have or ())
        need_obs = self._vector_scaler() is not None and not dev.pixel_shape and "obs" not in have
        need_ar = self._scales_actions_rewards(dev) and "act_rew" not in have
        mine = self._batches.get(dev.B)
        foreign = dev is not mine
        if foreign or ((need_obs or need_ar) and getattr(batch, "_transitions", None) is not None):
            if foreign:
                own = self.device_batch(dev.B)
            else:   # the caller's view of our own buffer must keep showing raw data: scale a private copy
                own = getattr(self, "_own_batch", None)
                if own is None or own.B != dev.B:
                    own = self._own_batch = self._make_batch(dev.B)
                    torch.cuda.current_stream(self._device).synchronize()
                    self._graphs_invalidate()
            self._lib.copy_d2d(own.dev.data_ptr(), dev.dev.data_ptr(), 4 * dev.nfloat, self._stream)
            if dev.pixel_shape:
                self._lib.copy_d2d(own.pix_dev.data_ptr(), dev.pix_dev.data_ptr(), 2 * dev.npix, self._stream)
            dev = own
        elif need_obs or need_ar:   # gathered straight into the impl's own buffer: scale in place, once
            batch.scaled = have | {"obs", "act_rew"}
        if need_obs:
            self._scale_observations(dev)
        if need_ar:
            self.scale_actions_rewards(dev)
        return dev

    def _upload(self, db: DeviceBatch) -> None:
        self._copy_h2d(db.dev.data_ptr(), db.host.data_ptr(), 4 * db.nfloat)
        if db.pixel_shape:
            self._copy_h2d(db.pix_dev.data_ptr(), db.pix_host.data_ptr(), 2 * db.npix)
        self._apply_scalers(db)

    def _apply_scalers(self, db: DeviceBatch):
        """TorchMiniBatch.__init__ (d3rlpy/torch_utility.py:179-185) on a host-staged batch: scaler.transform on
        obs / next_obs, action_scaler.transform on actions, reward_scaler.transform on rewards, as device kernels;
        pixel scaling is fused into the first conv load."""
        self._scale_observations(db)
        self.scale_actions_rewards(db)

    def _vector_scaler(self):
        sc = self._scaler
        return sc if sc is not None and hasattr(sc, "affine_f32") else None

    def _scale_observations(self, db: DeviceBatch) -> None:
        if self._vector_scaler() is None or db.pixel_shape:
            return
        mean, std, eps = self._scaler_params()
        self._lib.standardize(db.ptr("obs"), mean.data_ptr(), std.data_ptr(), eps, 2 * db.B, db.O, self._stream)

    def _scales_actions_rewards(self, db: DeviceBatch) -> bool:
        return (self._action_scaler is not None and not db.discrete) or self._reward_scaler is not None

    def scale_actions_rewards(self, db: DeviceBatch) -> None:
        """action_scaler.transform / reward_scaler.transform on the minibatch buffer (torch_utility.py:182-185)."""
        if self._action_scaler is not None and not db.discrete:
            mn, mx = self._action_scaler_params()
            self._lib.scale_actions(db.ptr("act"), mn.data_ptr(), mx.data_ptr(), db.B, db.A, self._stream)
        if self._reward_scaler is not None:
            lo, hi, sub, mul, div = self._reward_scaler.constants()
            self._lib.scale_rewards(db.ptr("rew"), db.B, lo, hi, sub, mul, div, self._stream)

    def _scaler_params(self):
        """(subtrahend, divisor, eps) of the observation scaler on the device: StandardScaler (mean, std, eps),
        MinMaxScaler (min, max - min, 0)."""
        if getattr(self, "_scaler_dev", None) is None:
            sub, div, eps = self._scaler.affine_f32()
            self._scaler_dev = (torch.tensor(sub, device=self._device), torch.tensor(div, device=self._device), eps)
            torch.cuda.synchronize(self._device)   # one-time upload on torch's stream; the kernels run on ours
        return self._scaler_dev

    def _action_scaler_params(self):
        if getattr(self, "_action_scaler_dev", None) is None:
            mn, mx = self._action_scaler.bounds_f32()
            assert mn.size == self._action_size, "action scaler bounds do not match the action size"
            self._action_scaler_dev = (torch.tensor(mn, device=self._device), torch.tensor(mx, device=self._device))
            torch.cuda.synchronize(self._device)
        return self._action_scaler_dev

    def unscale_actions(self, act: torch.Tensor) -> torch.Tensor:
        """action_scaler.reverse_transform on predicted / sampled actions (algos/torch/base.py:60-62,77-79), in place
        on the device before the read-back."""
        if self._action_scaler is not None:
            mn, mx = self._action_scaler_params()
            assert act.is_contiguous() and act.shape[-1] == self._action_size
            self._lib.unscale_actions(act.data_ptr(), mn.data_ptr(), mx.data_ptr(), act.numel() // self._action_size,
                                      self._action_size, self._stream)
        return act

    # ------------------------------------------------------------------ noise
    def noise_layout(self, B: int) -> Dict[str, tuple]:
        """name -> (kind, shape) in reference draw order; overridden per algorithm."""
        return {}

    def _noise_plan(self, B: int):
        layout = self.noise_layout(B)
        normals = [(k, s) for k, (kind, s) in layout.items() if kind == "normal"]
        uniforms = [(k, s) for k, (kind, s) in layout.items() if kind == "uniform"]
        off, plan = 0, {}
        for k, s in normals + uniforms:
            plan[k] = (off, s)
            off += int(np.prod(s))
        n_normal = sum(int(np.prod(s)) for _, s in normals)
        return plan, n_normal, off - n_normal, list(layout.keys())

    def _noise_arena(self, B: int) -> torch.Tensor:
        """One noise arena per batch size (pointers captured in that size's graphs never move)."""
        t = self._noise_by_b.get(B)
        if t is None:
            _, n_norm, n_uni, _ = self._noise_plan(B)
            t = torch.zeros(max(4, _align4(n_norm + n_uni)), dtype=torch.float32, device=self._device)
            self._noise_by_b[B] = t
            torch.cuda.current_stream(self._device).synchronize()
        return t

    def noise_view(self, name: str, B: int) -> torch.Tensor:
        plan = self._noise_plan(B)[0]
        off, shape = plan[name]
        return self._noise_arena(B)[off:off + int(np.prod(shape))].view(shape)

    def inject_noise(self, tensors: List[torch.Tensor], B: int, names: Optional[List[str]] = None):
        """Parity mode: replay recorded draws (in reference draw order) instead of Philox."""
        order = names or self._noise_plan(B)[3]
        assert len(order) == len(tensors), (order, [tuple(t.shape) for t in tensors])
        for name, t in zip(order, tensors):
            v = self.noise_view(name, B)
            assert v.numel() == t.numel(), (name, tuple(v.shape), tuple(t.shape))
            v.copy_(t.to(self._device, torch.float32).reshape(v.shape))
        torch.cuda.current_stream(self._device).synchronize()
        if not self._noise_injected:
            self._noise_injected = True
            self._graphs_invalidate()

    def clear_injected_noise(self):
        if self._noise_injected:
            self._noise_injected = False
            self._graphs_invalidate()

    def fill_noise(self, B: int):
        plan, n_norm, n_uni, _ = self._noise_plan(B)
        if n_norm + n_uni == 0 or self._noise_injected:
            return
        # per-rank Philox key: ranks draw independent noise for their own rows
        seed = (self._seed + 0x9E3779B97F4A7C15 * self.rank) & 0xFFFFFFFFFFFFFFFF
        self._lib.noise_fill(self._noise_arena(B).data_ptr(), n_norm, n_uni, seed, self.counter_ptr(0), self._stream)

    # ------------------------------------------------------------------ graphs
    def _graphs_invalidate(self):
        for g in self._graphs.values():
            self._lib.graph_destroy(g)
        self._graphs.clear()
        self._graph_nodes.clear()

    def run_program(self, key: tuple, program) -> None:
        """Runs `program()` (a sequence of launches on self._stream) — captured once per `key` as a CUDA
        graph and replayed afterwards."""
        pend = getattr(self, "_pending_upload", None)
        self._pending_upload = None
        inner = program

        def program():  # noqa: F811 - host-batch upload in front, metric read-back behind, all in one graph
            if pend is not None:
                self._upload(pend)
            inner()
            if pend is not None:  # host caller: it will read the metrics right away
                self._copy_d2h(self._slots_host.data_ptr(), self._slots.data_ptr(), 4 * 64)

        key = tuple(key) + (pend is not None, id(self._batch))
        self._metrics_on_host = pend is not None
        if not self.use_graph:
            program()
            return
        g = self._graphs.get(key)
        if g is None:
            import ctypes

            # allocation-only pass: every workspace the program touches is created now, because
            # allocating inside stream capture is illegal
            self._lib.dry = True
            try:
                program()
            finally:
                self._lib.dry = False
            torch.cuda.synchronize(self._device)
            self._lib.graph_begin(self._stream)
            try:
                program()
            finally:
                exec_ = ctypes.c_void_p()
                nodes = ctypes.c_int()
                self._lib.graph_end(self._stream, ctypes.byref(exec_), ctypes.byref(nodes))
            g = exec_.value
            self._graphs[key] = g
            self._graph_nodes[key] = nodes.value
        self._lib.graph_launch(g, self._stream)

