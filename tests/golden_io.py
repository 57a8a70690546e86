"""Helpers to read tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import os
from collections import OrderedDict

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Case:
    def __init__(self, z, name):
        self.z, self.name = z, name
        keys = [str(k) for k in z[f"{name}/cfg_keys"]]
        self.cfg = dict(zip(keys, z[f"{name}/cfg_vals"]))
        self.steps = int(self.cfg["steps"])
        self.metric_keys = [str(k) for k in z[f"{name}/metric_keys"]]
        self.metrics = z[f"{name}/metrics"]

    def group(self, kind, grp):
        pre = f"{self.name}/{kind}/{grp}/"
        return OrderedDict((k[len(pre):], torch.tensor(self.z[k])) for k in self.z.files if k.startswith(pre))

    def batch(self, s):
        pre = f"{self.name}/batch{s}/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def noise(self, s):
        pre = f"{self.name}/noise{s}/"
        n = len([k for k in self.z.files if k.startswith(pre)])
        return [torch.tensor(self.z[f"{pre}{j}"]) for j in range(n)]

    def step_metrics(self, s):
        return {k: float(v) for k, v in zip(self.metric_keys, self.metrics[s]) if not np.isnan(v)}


def load_update():
    return np.load(os.path.join(GOLDEN, "update.npz"))


def load_sampler():
    return np.load(os.path.join(GOLDEN, "sampler.npz"))


def load_siblings():
    """SAC / TD3 fixtures (tests/golden/make_golden_siblings.py)."""
    return np.load(os.path.join(GOLDEN, "update_siblings.npz"))


def load_qr():
    """DiscreteCQL / DQN with the quantile-regression Q head (tests/golden/make_golden_qr.py)."""
    return np.load(os.path.join(GOLDEN, "update_qr.npz"))


def load_online():
    """Online ReplayBuffer scripts and sampled minibatches (tests/golden/make_golden_online.py)."""
    return np.load(os.path.join(GOLDEN, "online.npz"))


def load_scalers():
    """Observation / action / reward scaler fixtures (tests/golden/make_golden_scalers.py)."""
    return np.load(os.path.join(GOLDEN, "scalers.npz"))


def load_awac():
    """AWAC fixtures (tests/golden/make_golden_awac.py); oracle-only until the CUDA path exists."""
    return np.load(os.path.join(GOLDEN, "update_awac.npz"))
