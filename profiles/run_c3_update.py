"""A few eager c3 BCQ updates (bf16 mode) for ncu launch lists: python profiles/run_c3_update.py"""
import os
import sys
from types import SimpleNamespace

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3rlpy_b200.algos import BCQ  # noqa: E402

O, A, B, N = 17, 6, 256, 100
algo = BCQ(actor_encoder_factory=[400, 300], critic_encoder_factory=[400, 300], imitator_encoder_factory=[750, 750],
           batch_size=B, n_action_samples=N, precision="bf16")
algo.create_impl((O,), A)
algo.impl.use_graph = False
rs = np.random.RandomState(0)
batch = SimpleNamespace(observations=rs.randn(B, O).astype(np.float32),
                        actions=rs.uniform(-1, 1, (B, A)).astype(np.float32), rewards=rs.randn(B, 1).astype(np.float32),
                        next_observations=rs.randn(B, O).astype(np.float32), terminals=np.zeros((B, 1), np.float32),
                        n_steps=np.ones((B, 1), np.float32))
for _ in range(3):
    m = algo.update(batch)
print({k: float(v) for k, v in m.items()})
