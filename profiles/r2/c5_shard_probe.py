"""c5 (ant-shaped CQL, 10 critics) at the per-rank batch of the 8-GPU strong-scaling point (1024 rows) on ONE GPU:
graph-timed update and per-family device times — what a rank computes, without any exchange."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402

w = dict(bench.WORKLOADS["c5"])
for B in (1024, 8192):
    w["batch"] = B
    r = bench.Runner(w, 1, 0, 0, "bf16", False)
    K, W = 60, 10
    idx = r.indices(K + W)
    for i in range(W):
        r.step_device(idx[i])
    r.barrier()
    st = torch.cuda.ExternalStream(r.impl._stream)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for i in range(W, W + K):
        r.step_device(idx[i])
    b.record(st)
    r.barrier()
    print(f"c5 shape, batch {B}, one GPU: {a.elapsed_time(b) / K * 1e3:.1f} us per update", flush=True)
    rs = np.random.RandomState(0)
    hb = SimpleNamespace(observations=rs.randn(B, w["obs"]).astype(np.float32),
                         actions=rs.uniform(-1, 1, (B, w["act"])).astype(np.float32),
                         rewards=rs.randn(B, 1).astype(np.float32),
                         next_observations=rs.randn(B, w["obs"]).astype(np.float32),
                         terminals=np.zeros((B, 1), np.float32), n_steps=np.ones((B, 1), np.float32))
    fams, _ = bench.kernel_profile(r.algo, hb, n_iter=3)
    for k, v in fams.items():
        print(f"   {k:24s} n={v['launches_per_update']:4.1f} us={v['us_per_update']:8.1f}")
    del r
    torch.cuda.empty_cache()
