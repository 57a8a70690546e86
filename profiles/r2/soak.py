"""Soak run: tens of thousands of graph-replayed updates per configuration (device-resident batches, Philox noise),
metrics checked for finiteness every 1000 updates — rare pipeline hangs (bounded mbarrier / flag spins trap) or
corruption would show up here."""
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from d3rlpy_b200.algos import BCQ, DiscreteCQL  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for wname in ("c2", "c1"):
    for prec in ("bf16", "fp32"):
        r = bench.Runner(bench.WORKLOADS[wname], 1, 0, 0, prec, False)
        idx = r.indices(256)
        t0 = time.perf_counter()
        for i in range(N):
            r.step_device(idx[i % 256])
            if i % 1000 == 999:
                m = r.impl.read_slots()[:8]
                assert np.all(np.isfinite(m)), (wname, prec, i, m)
        torch.cuda.synchronize()
        print(f"{wname} {prec}: {N} updates ok, {(time.perf_counter() - t0) / N * 1e6:.1f} us per update incl. host,"
              f" last metrics {r.impl.read_slots()[:6]}", flush=True)
        del r
rs = np.random.RandomState(0)
pix = [SimpleNamespace(observations=rs.randint(0, 256, (32, 4, 84, 84)).astype(np.uint8),
                       actions=rs.randint(0, 4, 32).astype(np.int32), rewards=(rs.rand(32, 1) < 0.1).astype(np.float32),
                       next_observations=rs.randint(0, 256, (32, 4, 84, 84)).astype(np.uint8),
                       terminals=(rs.rand(32, 1) < 0.01).astype(np.float32), n_steps=np.ones((32, 1), np.float32))
       for _ in range(4)]
vec = [SimpleNamespace(observations=rs.randn(256, 17).astype(np.float32),
                       actions=rs.uniform(-1, 1, (256, 6)).astype(np.float32), rewards=rs.randn(256, 1).astype(np.float32),
                       next_observations=rs.randn(256, 17).astype(np.float32),
                       terminals=(rs.rand(256, 1) < 0.01).astype(np.float32), n_steps=np.ones((256, 1), np.float32))
       for _ in range(4)]
for prec in ("bf16", "fp32"):
    for name, algo, batches, shape, act in (
            ("c4", DiscreteCQL(batch_size=32, n_frames=4, scaler="pixel", precision=prec, target_update_interval=100), pix,
             (4, 84, 84), 4),
            ("c3", BCQ(actor_encoder_factory=[400, 300], critic_encoder_factory=[400, 300],
                       imitator_encoder_factory=[750, 750], batch_size=256, n_action_samples=100, precision=prec), vec,
             (17,), 6)):
        algo.create_impl(shape, act)
        n = max(1000, N // 10)
        for i in range(n):
            m = algo.update(batches[i % 4])
            if i % 500 == 499:
                assert all(np.isfinite(float(v)) for v in m.values()), (name, prec, i, m)
        print(f"{name} {prec}: {n} updates ok, last metrics {m}", flush=True)
        del algo
        torch.cuda.empty_cache()
