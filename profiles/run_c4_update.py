"""A few eager c4 DiscreteCQL updates (Nature DQN encoder, 4x84x84 uint8 frames, batch 32) for ncu launch lists:
python profiles/run_c4_update.py [precision]"""
import os
import sys
from types import SimpleNamespace

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3rlpy_b200.algos import DiscreteCQL  # noqa: E402

B, A = 32, 6
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
algo = DiscreteCQL(batch_size=B, precision=prec, scaler="pixel")
algo.create_impl((4, 84, 84), A)
algo.impl.use_graph = False
rs = np.random.RandomState(0)
batch = SimpleNamespace(observations=rs.randint(0, 256, (B, 4, 84, 84)).astype(np.uint8),
                        actions=rs.randint(0, A, B).astype(np.int32), rewards=rs.randn(B, 1).astype(np.float32),
                        next_observations=rs.randint(0, 256, (B, 4, 84, 84)).astype(np.uint8),
                        terminals=np.zeros((B, 1), np.float32), n_steps=np.ones((B, 1), np.float32))
for _ in range(3):
    m = algo.update(batch)
print({k: float(v) for k, v in m.items()})
