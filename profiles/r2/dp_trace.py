"""In-graph timing of the peer-exchange kernels of the sharded c2 CQL update (torchrun, N >= 2): how long each
rendezvous waits, how long block 0 of each exchange kernel runs, and the gaps between them (d3b_peer_set_trace).
usage: python -m torch.distributed.run --nproc-per-node N profiles/r2/dp_trace.py [c2|c5]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wname = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.WORKLOADS[wname]
r = bench.Runner(w, world, rank, local, "bf16", wname == "c5")
L = r.impl._lib
K, W = 300, 20
idx = r.indices(K + W)
for i in range(W):
    r.step_device(idx[i])
r.barrier()
trace = torch.zeros(4 * 64 * 4, dtype=torch.int64, device=r.dev)
L.peer_set_trace(trace.data_ptr())
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.ExternalStream(r.impl._stream)
a.record(st)
for i in range(W, W + K):
    r.step_device(idx[i])
b.record(st)
r.barrier()
L.peer_set_trace(None)
us = a.elapsed_time(b) / K * 1e3
t = trace.cpu().numpy().reshape(4, 64, 4)
names = ["peer_wait_zero", "dp_scalar_steps", "adam critic", "adam actor"]
mhz = 1965.0
out = [f"rank {rank}/{world} {wname}: {us:.1f} us per update"]
ent = {}
for k in range(4):
    ok = t[k, :, 3] > 0
    if not ok.any():
        continue
    ent[k] = t[k, :, 0].astype(np.float64)
    out.append(f"  {names[k]:16s} wait {t[k, ok, 1].mean() / mhz:6.1f} us   block 0 total {t[k, ok, 2].mean() / mhz:6.1f} us")
if len(ent) == 4:
    order = [0, 1, 2, 3]
    for i in range(4):
        k0, k1 = order[i], order[(i + 1) % 4]
        d = ent[k1] - ent[k0] if i < 3 else np.roll(ent[k1], -1) - ent[k0]
        d = d[(d > 0) & (d < 1e6)]
        out.append(f"  entry {names[k0]:16s} -> entry {names[k1]:16s} {d.mean() / 1e3:6.1f} us")
print("\n".join(out), flush=True)
dist.destroy_process_group()
