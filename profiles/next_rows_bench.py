"""Timings for the section-8f rows widened in round 1 (quantile-regression DiscreteCQL at the Atari reproduction shape,
IQL / DDPG at the c1 shape, the HBM online ReplayBuffer) through the public API, with the CPU oracle port beside them.
Usage (GPU box): python profiles/next_rows_bench.py > profiles/r1_next_rows.json"""
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import update as ou  # noqa: E402  (CPU baseline leg only)


def time_updates(algo, batches, n=200, warm=20):
    for i in range(warm):
        algo.update(batches[i % len(batches)])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        algo.update(batches[i % len(batches)])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def time_oracle(orc, batches, scaler=None, budget=6.0):
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    obs = [ou.Batch(b, scaler) for b in batches]
    orc.update(obs[0], ou.Noise(seed=0))
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget:
        orc.update(obs[n % len(obs)], ou.Noise(seed=n))
        n += 1
    return (time.perf_counter() - t0) / n * 1e6


out = {}
rs = np.random.RandomState(0)

# ---- QR DiscreteCQL, Atari reproduction shape (reproductions/offline/discrete_cql.py: 200 quantiles), batch 32
from d3rlpy_b200.algos import DDPG, IQL, DiscreteCQL, QRQFunctionFactory  # noqa: E402

B, A, NQ = 32, 4, 200
pix = [dict(observations=rs.randint(0, 256, (B, 4, 84, 84)).astype(np.uint8),
            next_observations=rs.randint(0, 256, (B, 4, 84, 84)).astype(np.uint8),
            actions=rs.randint(A, size=B).astype(np.int32), rewards=(rs.rand(B, 1) < 0.1).astype(np.float32),
            terminals=np.zeros((B, 1), np.float32), n_steps=np.ones((B, 1), np.float32)) for _ in range(4)]
row = {}
for precision in ("bf16", "fp32"):
    algo = DiscreteCQL(batch_size=B, n_frames=4, scaler="pixel", q_func_factory=QRQFunctionFactory(n_quantiles=NQ),
                       precision=precision)
    algo.create_impl((4, 84, 84), A)
    row[f"us_per_update_{precision}"] = time_updates(algo, [SimpleNamespace(**b) for b in pix])
row["us_per_update_cpu_oracle"] = time_oracle(ou.DiscreteCQL((4, 84, 84), A, n_quantiles=NQ), pix, ou.pixel_scaler())
out["DiscreteCQL + QR(200), Nature DQN, batch 32 uint8 4x84x84 (host batch in, metric out)"] = row

# ---- IQL and DDPG at the c1 shape (obs 11, act 3, batch 256, 256x256)
O, A, B = 11, 3, 256
vec = [dict(observations=rs.randn(B, O).astype(np.float32), actions=rs.uniform(-1, 1, (B, A)).astype(np.float32),
            rewards=rs.randn(B, 1).astype(np.float32), next_observations=rs.randn(B, O).astype(np.float32),
            terminals=np.zeros((B, 1), np.float32), n_steps=np.ones((B, 1), np.float32)) for _ in range(4)]
for name, cls, orc in (("IQL", IQL, ou.IQL(O, A)), ("DDPG", DDPG, ou.DDPG(O, A))):
    row = {}
    for precision in ("bf16", "fp32"):
        algo = cls(batch_size=B, precision=precision)
        algo.create_impl((O,), A)
        row[f"us_per_update_{precision}"] = time_updates(algo, [SimpleNamespace(**b) for b in vec])
    row["us_per_update_cpu_oracle"] = time_oracle(orc, vec)
    out[f"{name}, obs 11 / act 3, batch 256, 256x256 (host batch in, metrics out)"] = row

# ---- online ReplayBuffer: append rate, and sample(256) from a full 1M-transition buffer (c2 shapes)
from d3rlpy_b200.online import ReplayBuffer  # noqa: E402

O, A, N = 17, 6, 1_000_000
env = SimpleNamespace(observation_space=SimpleNamespace(shape=(O,)), action_space=SimpleNamespace(shape=(A,)))
buf = ReplayBuffer(N, env=env)
obs, act = rs.randn(4096, O).astype(np.float32), rs.uniform(-1, 1, (4096, A)).astype(np.float32)
t0 = time.perf_counter()
for i in range(N + 5000):
    last = i % 1000 == 999
    buf.append(obs[i & 4095], act[i & 4095], 0.1, 1.0 if last else 0.0, clip_episode=last)
buf.flush()
append_s = time.perf_counter() - t0
np.random.seed(0)
for _ in range(20):
    buf.sample(256)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(500):
    buf.sample(256)
torch.cuda.synchronize()
sample_us = (time.perf_counter() - t0) / 500 * 1e6
t0 = time.perf_counter()
for i in range(2000):   # the online loop's pattern: one append, one sample
    buf.append(obs[i & 4095], act[i & 4095], 0.1, 0.0, clip_episode=False)
    buf.sample(256)
torch.cuda.synchronize()
out["online ReplayBuffer, 1M transitions, obs 17 / act 6"] = {
    "append_us_per_step": append_s / (N + 5000) * 1e6, "sample_256_us": sample_us,
    "append_plus_sample_256_us": (time.perf_counter() - t0) / 2000 * 1e6, "len": len(buf),
    "reference_cpu_in_build_container": "see profiles/r1_next_rows.md"}
print(json.dumps(out, indent=1))
