"""us/update of every BASELINE.json config through the public API (algo.update on a host numpy batch: pinned H2D +
one CUDA-graph replay + metric D2H), beside the CPU oracle port on the same host.  Supplementary to bench.py, which
measures the headline config c2.   python profiles/all_configs_bench.py > profiles/r1_all_configs.json"""
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3rlpy_b200.algos import BCQ, CQL, DiscreteCQL, TD3PlusBC  # noqa: E402
from oracle import update as ou  # noqa: E402

rs = np.random.RandomState(0)


def vec_batch(B, O, A):
    return dict(observations=rs.randn(B, O).astype(np.float32), actions=rs.uniform(-1, 1, (B, A)).astype(np.float32),
                rewards=rs.randn(B, 1).astype(np.float32), next_observations=rs.randn(B, O).astype(np.float32),
                terminals=(rs.rand(B, 1) < 0.01).astype(np.float32), n_steps=np.ones((B, 1), np.float32))


def pix_batch(B, A):
    return dict(observations=rs.randint(0, 256, (B, 4, 84, 84)).astype(np.uint8),
                actions=rs.randint(0, A, B).astype(np.int32), rewards=(rs.rand(B, 1) < 0.1).astype(np.float32),
                next_observations=rs.randint(0, 256, (B, 4, 84, 84)).astype(np.uint8),
                terminals=(rs.rand(B, 1) < 0.01).astype(np.float32), n_steps=np.ones((B, 1), np.float32))


def time_gpu(algo, batches, n=300, warm=20):
    hb = [SimpleNamespace(**b) for b in batches]
    for i in range(warm):
        algo.update(hb[i % len(hb)])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        algo.update(hb[i % len(hb)])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def time_cpu(orc, batches, scaler=None, budget=6.0):
    torch.set_num_threads(os.cpu_count() or 1)
    noise = ou.Noise(seed=0)
    orc.update(ou.Batch(batches[0], scaler), noise)
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget:
        orc.update(ou.Batch(batches[n % len(batches)], scaler), noise)
        n += 1
    return (time.perf_counter() - t0) / n * 1e6, n


out = {}
H3 = [256, 256, 256]
cfgs = []
# c1 TD3+BC
b = [vec_batch(256, 11, 3) for _ in range(4)]
for prec in ("fp32", "bf16"):
    a = TD3PlusBC(scaler=None, precision=prec)
    a.create_impl((11,), 3)
    out[f"c1 TD3+BC B256 ({prec})"] = {"us_per_update": time_gpu(a, b)}
us, n = time_cpu(ou.TD3PlusBC(11, 3), b)
out["c1 TD3+BC B256 (cpu oracle)"] = {"us_per_update": us, "updates_timed": n, "threads": os.cpu_count()}
# c2 CQL
b = [vec_batch(256, 17, 6) for _ in range(4)]
for prec in ("fp32", "bf16"):
    a = CQL(actor_encoder_factory=H3, critic_encoder_factory=H3, n_action_samples=10, precision=prec)
    a.create_impl((17,), 6)
    out[f"c2 CQL B256 N10 ({prec})"] = {"us_per_update": time_gpu(a, b)}
us, n = time_cpu(ou.CQL(17, 6, hidden=H3, n_action_samples=10), b)
out["c2 CQL B256 N10 (cpu oracle)"] = {"us_per_update": us, "updates_timed": n, "threads": os.cpu_count()}
# c3 BCQ (B256 and the script default B100)
for B in (256, 100):
    b = [vec_batch(B, 17, 6) for _ in range(4)]
    for prec in ("fp32", "bf16"):
        a = BCQ(actor_encoder_factory=[400, 300], critic_encoder_factory=[400, 300], imitator_encoder_factory=[750, 750],
                batch_size=B, n_action_samples=100, precision=prec)
        a.create_impl((17,), 6)
        out[f"c3 BCQ B{B} N100 ({prec})"] = {"us_per_update": time_gpu(a, b, n=100, warm=10)}
    us, n = time_cpu(ou.BCQ(17, 6, n_action_samples=100), b)
    out[f"c3 BCQ B{B} N100 (cpu oracle)"] = {"us_per_update": us, "updates_timed": n, "threads": os.cpu_count()}
# c4 DiscreteCQL pixels
b = [pix_batch(32, 4) for _ in range(4)]
for prec in ("fp32", "bf16"):
    a = DiscreteCQL(batch_size=32, n_frames=4, scaler="pixel", precision=prec)
    a.create_impl((4, 84, 84), 4)
    out[f"c4 DiscreteCQL 4x84x84 B32 ({prec})"] = {"us_per_update": time_gpu(a, b, n=100, warm=10)}
us, n = time_cpu(ou.DiscreteCQL((4, 84, 84), 4), b, scaler=ou.pixel_scaler())
out["c4 DiscreteCQL 4x84x84 B32 (cpu oracle)"] = {"us_per_update": us, "updates_timed": n, "threads": os.cpu_count()}
# fit(): sampling (numpy index stream) + device gather + update, c2 shapes on a 1M-step replay
from d3rlpy_b200.dataset import MDPDataset  # noqa: E402

S = 1_000_000
ds = MDPDataset(rs.randn(S, 17).astype(np.float32), rs.uniform(-1, 1, (S, 6)).astype(np.float32),
                rs.randn(S).astype(np.float32), (np.arange(S) % 1000 == 999).astype(np.float32))
a = CQL(actor_encoder_factory=H3, critic_encoder_factory=H3, n_action_samples=10, precision="bf16")
a.fit(ds, n_steps=200, n_steps_per_epoch=100, seed=0)
t0 = time.perf_counter()
a.fit(ds, n_steps=3000, n_steps_per_epoch=1000, seed=1)
out["c2 CQL fit() sample+gather+update (bf16)"] = {"us_per_update": (time.perf_counter() - t0) / 3000 * 1e6}
for k, v in out.items():
    v["updates_per_s"] = 1e6 / v["us_per_update"]
print(json.dumps(out, indent=1))
