"""Data-parallel plumbing: one process per GPU, the minibatch sharded by rows, one sum all-reduce per
optimizer step.  The reference has no distributed path; this is new.

* `shard_rows(B, world, rank)` — rank r owns rows [r*B/W, (r+1)*B/W) of the SAME global index vector
  (all ranks draw identical indices from a shared seed; the replay buffer is replicated per GPU).
* `shard_noise(t, kind, B, N, world, rank)` — slices injected global noise so that a sharded update
  reproduces the single-GPU update bit-for-bit up to summation order (row = b*N + k keeps a sample's N
  actions on one rank).
* `init(world, rank)` / `allreduce_sum(tensor, stream)` — NCCL communicator owned by the C-ABI library
  (d3b_comm_*), rendezvous of the 128-byte unique id through torch.distributed (any backend).
"""
from __future__ import annotations

import ctypes
import glob
import os
from typing import Optional, Tuple

import torch

from ._lib import D3BError, lib

_COMM: Optional[int] = None
_WORLD = 1
_RANK = 0


def shard_rows(batch: int, world: int, rank: int) -> Tuple[int, int]:
    if batch % world:
        raise ValueError(f"global batch {batch} is not divisible by world size {world}")
    per = batch // world
    return rank * per, (rank + 1) * per


def shard_noise(t: torch.Tensor, shape_kind: str, batch: int, n: int, world: int, rank: int) -> torch.Tensor:
    """shape_kind: "B*" (B, ...) rows;  "NB*" (N, B, ...) as Normal.rsample((N,));  "BN*" (B*N, ...) with
    row = b*N + k (cql_impl.py:155-161,176-186)."""
    lo, hi = shard_rows(batch, world, rank)
    if shape_kind == "B*":
        return t[lo:hi].contiguous()
    if shape_kind == "NB*":
        return t[:, lo:hi].contiguous()
    if shape_kind == "BN*":
        return t[lo * n:hi * n].contiguous()
    raise ValueError(shape_kind)


def _find_nccl() -> str:
    try:
        import nvidia.nccl  # the copy torch itself loaded

        cands = glob.glob(os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so*"))
        if cands:
            return sorted(cands)[0]
    except Exception:  # noqa: BLE001
        pass
    return "libnccl.so.2"


def init(world_size: int, rank: int) -> None:
    """Creates the library-owned NCCL communicator.  torch.distributed must already be initialised (it is
    only used to broadcast the unique id)."""
    global _COMM, _WORLD, _RANK
    if _COMM is not None or world_size <= 1:
        _WORLD, _RANK = max(world_size, 1), rank
        return
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        raise D3BError("parallel.init needs an initialised torch.distributed process group for the rendezvous")
    L = lib()
    L.comm_load(_find_nccl().encode())
    buf = (ctypes.c_uint8 * 128)()
    if rank == 0:
        L.comm_unique_id(ctypes.addressof(buf))
    obj = [bytes(buf)]
    dist.broadcast_object_list(obj, src=0)
    ctypes.memmove(ctypes.addressof(buf), obj[0], 128)
    comm = ctypes.c_void_p()
    L.comm_init(ctypes.addressof(buf), world_size, rank, ctypes.byref(comm))
    _COMM, _WORLD, _RANK = comm.value, world_size, rank


def world() -> Tuple[int, int]:
    return _WORLD, _RANK


def allreduce_sum(t: torch.Tensor, stream) -> None:
    """In-place sum over ranks of a contiguous fp32 device tensor, on `stream` (graph-capturable)."""
    if _WORLD <= 1:
        return
    if _COMM is None:
        raise D3BError("parallel.allreduce_sum before parallel.init")
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    cs = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
    lib().allreduce_sum(_COMM, t.data_ptr(), t.numel(), cs)


def destroy() -> None:
    global _COMM
    if _COMM is not None:
        lib().comm_destroy(_COMM)
        _COMM = None


# ----------------------------------------------------------------------------------- NVLink peer-memory exchange
class PeerExchange:
    """Maps device buffers of every rank of the box into this process (CUDA IPC) so that kernels can read the
    other ranks' gradient arenas over NVLink directly (csrc/comm.cu: adam_allreduce / peer_allreduce_small).
    All methods are collective: every rank must call them in the same order."""

    N_FLAGS = 32          # flags per rank: {ready, done} pairs; channels 0-3 small vectors, 4.. gradient arenas
    MAX_RANKS = 8         # every flag has one int32 slot per WRITER rank (pushed flags, csrc/comm.cu)
    XCHG_FLOATS = 4 * 2 * 8 * 16   # [channel][epoch parity][WRITER rank][16 floats]: partial sums are pushed too

    def __init__(self, world_size: int, rank: int, device):
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise D3BError("PeerExchange needs an initialised torch.distributed process group for the rendezvous")
        self.world, self.rank = world_size, rank
        self.flags = torch.zeros(self.N_FLAGS * self.MAX_RANKS, dtype=torch.int32, device=device)
        self.xchg = torch.zeros(self.XCHG_FLOATS, dtype=torch.float32, device=device)
        self.counters = torch.zeros(16, dtype=torch.int32, device=device)  # block counters of the fused Adam kernels
        torch.cuda.synchronize(device)
        self.flags_ptrs = self.register(self.flags)
        self.xchg_ptrs = self.register(self.xchg)
        self._next_arena = 0

    def register(self, t: torch.Tensor):
        """Returns a ctypes array of `world` device pointers: t of rank r as seen from this process."""
        import torch.distributed as dist

        L = lib()
        handle = (ctypes.c_uint8 * 64)()
        off = ctypes.c_int64()
        L.peer_export(t.data_ptr(), ctypes.addressof(handle), ctypes.byref(off))
        mine = (bytes(handle), int(off.value))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        ptrs = (ctypes.c_void_p * self.world)()
        for r, (h, o) in enumerate(everyone):
            if r == self.rank:
                ptrs[r] = t.data_ptr()
            else:
                buf = (ctypes.c_uint8 * 64).from_buffer_copy(h)
                out = ctypes.c_void_p()
                L.peer_import(ctypes.addressof(buf), o, ctypes.byref(out))
                ptrs[r] = out.value
        return ptrs

    def register_arena(self, grads: torch.Tensor):
        """Gradient arena -> (peer pointer table, flag index of its {ready, done, reduced} triple, block counter
        pointer, peer pointer table of the reduced-gradient buffers, second block counter pointer).  The reduced-
        gradient buffer (two-shot exchange, csrc/comm.cu) is allocated here, one per arena and rank."""
        i = self._next_arena
        self._next_arena += 1
        assert 8 + 3 * i + 2 < self.N_FLAGS and i < 8
        gred = torch.zeros_like(grads)
        torch.cuda.synchronize(grads.device)
        self._keep = getattr(self, "_keep", []) + [gred]
        # two-shot costs one more rendezvous inside the kernel (~14 us measured at W = 2) and saves (W-1) - 2 (W-1)/W
        # arena sizes of NVLink traffic: it pays for large arenas on many ranks (c5's 6.5 MB critic arena at W = 8: 45 MB
        # -> 11 MB per rank and update), not for c2's 1.1 MB.  D3B_TWO_SHOT=0 forces one-shot, =2 forces two-shot.
        mode = os.environ.get("D3B_TWO_SHOT", "1")
        remote_read_bytes = (self.world - 1) * grads.numel() * 4
        two_shot = mode == "2" or (mode != "0" and self.world > 2 and remote_read_bytes >= 16 * 1024 * 1024)
        return (self.register(grads), 8 + 3 * i, self.counters.data_ptr() + 4 * i,
                self.register(gred) if two_shot else None, self.counters.data_ptr() + 4 * (8 + i))


def new_peers(device=None) -> Optional[PeerExchange]:
    """A peer exchange for ONE data-parallel impl (None when world_size == 1 or D3B_PEER=0).  Collective: every rank
    must create its impls in the same order.  The flags are monotonically increasing update epochs of the owning impl,
    so two impls must never share a flag block (a second impl starting at epoch 0 would see every flag as already
    raised and read gradients that are not there yet)."""
    if _WORLD > 1 and os.environ.get("D3B_PEER", "1") != "0":
        return PeerExchange(_WORLD, _RANK, device)
    return None
