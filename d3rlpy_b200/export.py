"""`save_policy` (d3rlpy/algos/torch/base.py:86-126): the greedy policy as a TorchScript (`.pt`) or ONNX (`.onnx`)
function `observation -> action`, observation scaler included.

The reference traces its own `_predict_best_action`.  The accelerated path has no autograd modules to trace, so the
export rebuilds the same function in plain PyTorch from the `state_dict`s (reference key layout) and traces that —
deployment format, not the hot path: nothing here runs during training."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

NATURE_STRIDES = (4, 2, 1)


def _cpu(sd) -> Dict[str, torch.Tensor]:
    return {k: v.detach().to("cpu", torch.float32).clone() for k, v in sd.items()}


def _n_layers(p, prefix: str) -> int:
    n = 0
    while f"{prefix}_fcs.{n}.weight" in p:
        n += 1
    return n


def _mlp(p, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """VectorEncoder / VectorEncoderWithAction (models/torch/encoders.py:265-339), ReLU, no BN/dropout."""
    for i in range(_n_layers(p, prefix)):
        x = torch.relu(F.linear(x, p[f"{prefix}_fcs.{i}.weight"], p[f"{prefix}_fcs.{i}.bias"]))
    return x


def _encoder(p, prefix: str, x: torch.Tensor) -> torch.Tensor:
    if f"{prefix}_convs.0.weight" in p:  # PixelEncoder (encoders.py:130-162), Nature-DQN strides
        n = 0
        while f"{prefix}_convs.{n}.weight" in p:
            x = torch.relu(F.conv2d(x, p[f"{prefix}_convs.{n}.weight"], p[f"{prefix}_convs.{n}.bias"],
                                    stride=NATURE_STRIDES[n]))
            n += 1
        return torch.relu(F.linear(x.reshape(x.shape[0], -1), p[f"{prefix}_fc.weight"], p[f"{prefix}_fc.bias"]))
    return _mlp(p, prefix, x)


class GreedyPolicy(torch.nn.Module):
    """kind: "normal" (CQL/SAC: tanh(mu), policies.py:233-249), "deterministic" (TD3+BC, policies.py:57-59),
    "discrete" (DQN family: argmax of the member-mean Q, dqn_impl.py:131-133), "bcq" (bcq_impl.py:163-211)."""

    def __init__(self, kind: str, policy=None, q=None, imitator=None, scaler=None, n_action_samples: int = 100,
                 action_flexibility: float = 0.05, n_quantiles: int = 0, action_scaler=None):
        super().__init__()
        self.kind, self.n, self.flex = kind, n_action_samples, action_flexibility
        self.n_quantiles = n_quantiles  # > 0: QR members, value = mean over the quantiles (qr_q_function.py:44-48)
        self.pi = _cpu(policy) if policy is not None else {}
        self.q = _cpu(q) if q is not None else {}
        self.vae = _cpu(imitator) if imitator is not None else {}
        self.scaler_kind = getattr(scaler, "TYPE", scaler if isinstance(scaler, str) else None)
        if self.scaler_kind == "standard":
            self.mean = torch.as_tensor(scaler._mean, dtype=torch.float32)
            self.std = torch.as_tensor(scaler._std, dtype=torch.float32)
            self.eps = float(scaler._eps)
        if self.scaler_kind == "min_max":
            self.minimum = torch.as_tensor(scaler._minimum, dtype=torch.float32)
            self.maximum = torch.as_tensor(scaler._maximum, dtype=torch.float32)
        self.act_min = self.act_max = None
        if action_scaler is not None and kind != "discrete":
            self.act_min = torch.as_tensor(action_scaler._minimum, dtype=torch.float32)
            self.act_max = torch.as_tensor(action_scaler._maximum, dtype=torch.float32)

    def _scale(self, x: torch.Tensor) -> torch.Tensor:
        if self.scaler_kind == "standard":  # scalers.py:350-354
            return (x - self.mean) / (self.std + self.eps)
        if self.scaler_kind == "min_max":   # scalers.py:209-218
            return (x - self.minimum) / (self.maximum - self.minimum)
        if self.scaler_kind == "pixel":     # scalers.py:109-110
            return x.float() / 255.0
        return x

    def _q_members(self, x, action: Optional[torch.Tensor]):
        vals, i = [], 0
        while f"_q_funcs.{i}._fc.weight" in self.q:
            pre = f"_q_funcs.{i}._encoder."
            h = _encoder(self.q, pre, x if action is None else torch.cat([x, action], dim=1))
            v = F.linear(h, self.q[f"_q_funcs.{i}._fc.weight"], self.q[f"_q_funcs.{i}._fc.bias"])
            if self.n_quantiles > 0:
                v = v.view(v.shape[0], -1, self.n_quantiles).mean(dim=2)
            vals.append(v)
            i += 1
        return torch.stack(vals, 0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """`_func` of TorchImplBase.save_policy (algos/torch/base.py:91-100): scaler.transform, greedy action,
        action_scaler.reverse_transform."""
        action = self._greedy(self._scale(x))
        if self.act_min is not None:   # action_scalers.py:197-206
            action = ((self.act_max - self.act_min) * ((action + 1.0) / 2.0)) + self.act_min
        return action

    def _greedy(self, x: torch.Tensor) -> torch.Tensor:
        if self.kind == "normal":
            return torch.tanh(F.linear(_mlp(self.pi, "_encoder.", x), self.pi["_mu.weight"], self.pi["_mu.bias"]))
        if self.kind == "deterministic":
            return torch.tanh(F.linear(_mlp(self.pi, "_encoder.", x), self.pi["_fc.weight"], self.pi["_fc.bias"]))
        if self.kind == "discrete":
            return self._q_members(x, None).mean(0).argmax(dim=1)
        if self.kind == "bcq":
            b, a = x.shape[0], self.pi["_fc.weight"].shape[0]
            xr = x[:, None, :].expand(b, self.n, x.shape[1]).reshape(b * self.n, x.shape[1])
            latent = torch.randn(b * self.n, 2 * a).clamp(-0.5, 0.5)
            h = _mlp(self.vae, "_decoder_encoder.", torch.cat([xr, latent], dim=1))
            sampled = torch.tanh(F.linear(h, self.vae["_fc.weight"], self.vae["_fc.bias"]))
            h = _mlp(self.pi, "_encoder.", torch.cat([xr, sampled], dim=1))
            res = self.flex * torch.tanh(F.linear(h, self.pi["_fc.weight"], self.pi["_fc.bias"]))
            cand = (sampled + res).clamp(-1.0, 1.0)
            index = self._q_members(xr, cand)[0].view(b, self.n).argmax(dim=1)
            return cand.view(b, self.n, a)[torch.arange(b), index]
        raise ValueError(self.kind)


def save_policy(module: GreedyPolicy, observation_shape, fname: str) -> None:
    dummy = torch.rand(1, *observation_shape)
    with torch.no_grad():
        traced = torch.jit.trace(module, dummy, check_trace=False)
    if fname.endswith(".onnx"):
        torch.onnx.export(traced, dummy, fname, export_params=True, opset_version=11, input_names=["input_0"],
                          output_names=["output_0"])
    elif fname.endswith(".pt"):
        traced.save(fname)
    else:
        raise ValueError(f"invalid format type: {fname}. .pt and .onnx extensions are currently supported.")
