"""Runs a few eager (no CUDA graph) c2 CQL updates in bf16 mode so that ncu can capture individual launches of
the fused kernels:  ncu --set full -k regex:mlp_ ... python profiles/run_c2_update.py"""
import os
import sys
from types import SimpleNamespace

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3rlpy_b200.algos import CQL  # noqa: E402

O, A, B, N, H = 17, 6, 256, 10, [256, 256, 256]
PRECISION = sys.argv[2] if len(sys.argv) > 2 else "bf16"
algo = CQL(actor_encoder_factory=H, critic_encoder_factory=H, n_action_samples=N, precision=PRECISION)
algo.create_impl((O,), A)
algo.impl.use_graph = False
rs = np.random.RandomState(0)
batch = SimpleNamespace(observations=rs.randn(B, O).astype(np.float32),
                        actions=rs.uniform(-1, 1, (B, A)).astype(np.float32), rewards=rs.randn(B, 1).astype(np.float32),
                        next_observations=rs.randn(B, O).astype(np.float32), terminals=np.zeros((B, 1), np.float32),
                        n_steps=np.ones((B, 1), np.float32))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    m = algo.update(batch)
print({k: float(v) for k, v in m.items()})
