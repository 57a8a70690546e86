"""Stock CUDA-PyTorch denominator for the north-star target (">= 20x the reference's own CUDA-PyTorch updates/sec for
CQL at batch 256 on one B200").

/root/reference cannot travel to the GPU box, so what is timed is the oracle port of the reference update
(oracle/update.py: the same eager PyTorch fp32 operator sequence -- F.linear stacks per member, autograd,
torch.optim.Adam, per-parameter soft_sync, one host sync per returned loss like `loss.cpu().detach().numpy()`) with
its parameters, optimizer state and noise draws on cuda:0 and the numpy minibatch uploaded every update like
TorchMiniBatch does (torch_utility.py:146-177).  It is a reported baseline only; nothing here is on the product path.

    python profiles/torch_cuda_port_bench.py [--device cuda] [--steps 300] [--warmup 30] > profiles/r1_torch_cuda_port.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import update as ou  # noqa: E402


class DeviceNoise:
    """Fresh draws on the device, like torch.randn / uniform_ inside the reference's impl."""

    def __init__(self, device):
        self.device = device

    def normal(self, *shape):
        return torch.randn(*shape, device=self.device)

    def uniform(self, *shape):
        return torch.empty(*shape, device=self.device).uniform_(-1.0, 1.0)


def to_device(algo, device):
    """Moves every parameter dictionary of an oracle algorithm to `device` and rebuilds its optimizers there."""
    for name in ("q", "pi", "log_temp", "log_alpha", "targ_q", "targ_pi"):
        if hasattr(algo, name):
            d = getattr(algo, name)
            grad = not name.startswith("targ")
            for k in list(d):
                d[k] = d[k].detach().to(device).requires_grad_(grad)
    algo.critic_optim = ou.make_adam(algo.q, algo.critic_optim.param_groups[0]["lr"])
    algo.actor_optim = ou.make_adam(algo.pi, algo.actor_optim.param_groups[0]["lr"])
    if hasattr(algo, "temp_optim"):
        algo.temp_optim = ou.make_adam(algo.log_temp, algo.temp_lr)
        algo.alpha_optim = ou.make_adam(algo.log_alpha, algo.alpha_lr)


class DeviceBatch:
    """TorchMiniBatch: numpy -> device float32 tensors, every update."""

    def __init__(self, arrays, device):
        for k, v in arrays.items():
            setattr(self, k, torch.tensor(data=v, dtype=torch.float32, device=device))


def run(workload, device, steps, warmup, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    rs = np.random.RandomState(0)
    if workload == "c2":
        O, A, B = 17, 6, 256
        algo = ou.CQL(O, A, hidden=(256, 256, 256), n_action_samples=10, seed=0)
    else:  # c1
        O, A, B = 11, 3, 256
        algo = ou.TD3PlusBC(O, A, hidden=(256, 256), seed=0)
    to_device(algo, device)
    noise = DeviceNoise(device)
    batches = [dict(observations=rs.randn(B, O).astype(np.float32),
                    actions=rs.uniform(-1, 1, (B, A)).astype(np.float32), rewards=rs.randn(B, 1).astype(np.float32),
                    next_observations=rs.randn(B, O).astype(np.float32), terminals=np.zeros((B, 1), np.float32),
                    n_steps=np.ones((B, 1), np.float32)) for _ in range(8)]
    sync = (lambda: torch.cuda.synchronize()) if device.startswith("cuda") else (lambda: None)
    for i in range(warmup):
        algo.update(DeviceBatch(batches[i % 8], device), noise)
    sync()
    t = []
    for i in range(steps):
        t0 = time.perf_counter()
        m = algo.update(DeviceBatch(batches[i % 8], device), noise)
        sync()
        t.append(time.perf_counter() - t0)
    assert all(np.isfinite(v) for v in m.values()), m
    t = np.array(t)
    return {"workload": workload, "tf32": tf32, "updates_per_s": float(steps / t.sum()),
            "ms_per_update_p10_p50_p90": [float(np.percentile(t, q) * 1e3) for q in (10, 50, 90)]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    a = ap.parse_args()
    out = {"what": "oracle port of the reference update run by stock eager PyTorch on " + a.device
                   + " (parameters, Adam state and noise on the device; numpy minibatch uploaded per update)",
           "torch": torch.__version__, "host_threads": torch.get_num_threads(),
           "device_name": torch.cuda.get_device_name(0) if a.device.startswith("cuda") else "cpu", "runs": []}
    for workload in ("c2", "c1"):
        for tf32 in (False, True):
            out["runs"].append(run(workload, a.device, a.steps, a.warmup, tf32))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
