"""CPU oracle for the offline-RL update step.

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / ``--impl reference`` legs and the baseline legs of the
measurement scripts under profiles/ (as the thing the product is compared WITH);
the product path (d3rlpy_b200) never imports it and has no CPU fallback.

A plain-PyTorch (fp32, autograd, ``torch.optim.Adam``) restatement of what one
``algo.update(batch)`` of the reference computes, written functionally over
dicts of tensors whose keys equal the reference modules' ``state_dict`` keys so
weights can be exchanged with the reference and with the CUDA path.  Each
function cites the reference lines it follows (relative to /root/reference/).

Third-party arithmetic (GEMM, Adam, Normal, logsumexp) is torch 2.11.0 as in
the reference.  Pinned against the live reference by
tests/golden/update*.npz and scalers.npz (tests/golden/make_golden*.py) — see DESIGN.md §2.
AWAC, CRR, PLAS, BEAR, DiscreteBCQ, DiscreteSAC, TD3PlusRelation and BC / DiscreteBC are restated and
pinned here ahead of their CUDA paths.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- noise
class Noise:
    """Noise source.  ``injected`` is a list of tensors replayed in draw order
    (the order SURVEY.md §8c lists); otherwise fresh draws from ``generator``."""

    def __init__(self, injected: Optional[List[torch.Tensor]] = None, seed: int = 0):
        self.injected = list(injected) if injected is not None else None
        self.gen = torch.Generator().manual_seed(seed)
        self.log: List[torch.Tensor] = []

    def _next(self, shape, kind):
        if self.injected is not None:
            t = self.injected.pop(0)
            assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape), kind)
            return t.clone()
        if kind == "normal":
            t = torch.randn(shape, generator=self.gen)
        else:
            t = torch.rand(shape, generator=self.gen) * 2.0 - 1.0
        self.log.append(t.clone())
        return t

    def normal(self, *shape):
        return self._next(shape, "normal")

    def uniform(self, *shape):
        return self._next(shape, "uniform")


# --------------------------------------------------------------------------- parameter construction
def _linear_init(out_f: int, in_f: int, gen: torch.Generator):
    # nn.Linear default: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for W and b
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return w, b


def make_mlp(prefix: str, in_f: int, hidden: Sequence[int], gen) -> Params:
    """``_VectorEncoder`` Linear stack (d3rlpy/models/torch/encoders.py:236-263)."""
    p: Params = OrderedDict()
    for i, h in enumerate(hidden):
        w, b = _linear_init(h, in_f, gen)
        p[f"{prefix}_fcs.{i}.weight"], p[f"{prefix}_fcs.{i}.bias"] = w, b
        in_f = h
    return p


def make_head(prefix: str, out_f: int, in_f: int, gen) -> Params:
    w, b = _linear_init(out_f, in_f, gen)
    return OrderedDict([(f"{prefix}.weight", w), (f"{prefix}.bias", b)])


def make_critics(obs: int, act: int, hidden, n: int, gen, n_quantiles: Optional[int] = None) -> Params:
    """EnsembleContinuousQFunction of ContinuousMeanQFunction (builders.py:55-77); with ``n_quantiles`` of
    ContinuousQRQFunction (qr_q_function.py:101-110: head ``Linear(feature, n_quantiles)``)."""
    p: Params = OrderedDict()
    for i in range(n):
        p.update(make_mlp(f"_q_funcs.{i}._encoder.", obs + act, hidden, gen))
        p.update(make_head(f"_q_funcs.{i}._fc", n_quantiles or 1, hidden[-1], gen))
    return p


def make_squashed_normal_policy(obs: int, act: int, hidden, gen) -> Params:
    p = make_mlp("_encoder.", obs, hidden, gen)
    p.update(make_head("_mu", act, hidden[-1], gen))
    p.update(make_head("_logstd", act, hidden[-1], gen))
    return p


def make_deterministic_policy(obs: int, act: int, hidden, gen) -> Params:
    p = make_mlp("_encoder.", obs, hidden, gen)
    p.update(make_head("_fc", act, hidden[-1], gen))
    return p


def make_residual_policy(obs: int, act: int, hidden, gen) -> Params:
    p = make_mlp("_encoder.", obs + act, hidden, gen)
    p.update(make_head("_fc", act, hidden[-1], gen))
    return p


def make_cvae(obs: int, act: int, latent: int, hidden, gen) -> Params:
    """ConditionalVAE (d3rlpy/models/torch/imitators.py:13-61); registration order
    encoder_encoder, decoder_encoder, _mu, _logstd, _fc."""
    p = make_mlp("_encoder_encoder.", obs + act, hidden, gen)
    p.update(make_mlp("_decoder_encoder.", obs + latent, hidden, gen))
    p.update(make_head("_mu", latent, hidden[-1], gen))
    p.update(make_head("_logstd", latent, hidden[-1], gen))
    p.update(make_head("_fc", act, hidden[-1], gen))
    return p


NATURE_FILTERS = [(32, 8, 4), (64, 4, 2), (64, 3, 1)]


def conv_out_hw(h: int, w: int, filters=NATURE_FILTERS):
    for _, k, s in filters:
        h, w = (h - k) // s + 1, (w - k) // s + 1
    return h, w


def make_discrete_critics(obs_shape, act: int, n: int, gen, hidden=None, feature_size=512,
                          n_quantiles: Optional[int] = None) -> Params:
    """EnsembleDiscreteQFunction of DiscreteMeanQFunction over PixelEncoder
    (encoders.py:43-162) or VectorEncoder; with ``n_quantiles`` the members are DiscreteQRQFunction
    (qr_q_function.py:22-36: head ``Linear(feature, action_size * n_quantiles)``)."""
    p: Params = OrderedDict()
    for i in range(n):
        pre = f"_q_funcs.{i}._encoder."
        if len(obs_shape) == 3:
            c = obs_shape[0]
            for l, (oc, k, s) in enumerate(NATURE_FILTERS):
                bound = 1.0 / math.sqrt(c * k * k)
                p[f"{pre}_convs.{l}.weight"] = (torch.rand(oc, c, k, k, generator=gen) * 2 - 1) * bound
                p[f"{pre}_convs.{l}.bias"] = (torch.rand(oc, generator=gen) * 2 - 1) * bound
                c = oc
            hh, ww = conv_out_hw(obs_shape[1], obs_shape[2])
            w, b = _linear_init(feature_size, c * hh * ww, gen)
            p[f"{pre}_fc.weight"], p[f"{pre}_fc.bias"] = w, b
            feat = feature_size
        else:
            hidden = hidden or [256, 256]
            p.update(make_mlp(pre, obs_shape[0], hidden, gen))
            feat = hidden[-1]
        p.update(make_head(f"_q_funcs.{i}._fc", act * (n_quantiles or 1), feat, gen))
    return p


def clone_params(p: Params, requires_grad: bool = True) -> Params:
    return OrderedDict((k, v.detach().clone().requires_grad_(requires_grad)) for k, v in p.items())


# --------------------------------------------------------------------------- forward pieces
def _n_layers(p: Params, prefix: str) -> int:
    n = 0
    while f"{prefix}_fcs.{n}.weight" in p:
        n += 1
    return n


def mlp_forward(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """``_fc_encode`` with ReLU, no BN/dropout/dense (encoders.py:265-275)."""
    h = x
    for i in range(_n_layers(p, prefix)):
        h = torch.relu(F.linear(h, p[f"{prefix}_fcs.{i}.weight"], p[f"{prefix}_fcs.{i}.bias"]))
    return h


def pixel_forward(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """PixelEncoder.forward (encoders.py:106-140)."""
    h = x
    for l, (_, _, s) in enumerate(NATURE_FILTERS):
        h = torch.relu(F.conv2d(h, p[f"{prefix}_convs.{l}.weight"], p[f"{prefix}_convs.{l}.bias"], stride=s))
    return torch.relu(F.linear(h.reshape(h.shape[0], -1), p[f"{prefix}_fc.weight"], p[f"{prefix}_fc.bias"]))


def n_members(p: Params) -> int:
    n = 0
    while f"_q_funcs.{n}._fc.weight" in p:
        n += 1
    return n


def reduce_ensemble(y: torch.Tensor, reduction: str = "min", lam: float = 0.75) -> torch.Tensor:
    """``_reduce_ensemble`` (q_functions/ensemble_q_function.py:9-24)."""
    if reduction == "min":
        return y.min(dim=0).values
    if reduction == "max":
        return y.max(dim=0).values
    if reduction == "mean":
        return y.mean(dim=0)
    if reduction == "none":
        return y
    if reduction == "mix":
        return lam * y.min(dim=0).values + (1.0 - lam) * y.max(dim=0).values
    raise ValueError(reduction)


def q_continuous(p: Params, x, action, reduction="mean", lam=0.75) -> torch.Tensor:
    """EnsembleContinuousQFunction.forward (ensemble_q_function.py:163-170) over
    ContinuousMeanQFunction.forward (mean_q_function.py:71-72) and
    VectorEncoderWithAction.forward (encoders.py:328-339)."""
    vals = quantiles_continuous(p, x, action).mean(dim=2, keepdim=True)  # QR: mean over quantiles (:118-122)
    return reduce_ensemble(vals, reduction, lam)


def quantiles_continuous(p: Params, x, action) -> torch.Tensor:
    """Head outputs of every member, (E, B, n): n = 1 for ContinuousMeanQFunction, n_quantiles for
    ContinuousQRQFunction._compute_quantiles (qr_q_function.py:112-116)."""
    xa = torch.cat([x, action], dim=1)
    vals = []
    for i in range(n_members(p)):
        h = mlp_forward(p, f"_q_funcs.{i}._encoder.", xa)
        q = F.linear(h, p[f"_q_funcs.{i}._fc.weight"], p[f"_q_funcs.{i}._fc.bias"])
        vals.append(q.view(1, x.shape[0], -1))
    return torch.cat(vals, dim=0)


def q_target_continuous(p: Params, x, action) -> torch.Tensor:
    """EnsembleContinuousQFunction.compute_target(x, action, "min") (ensemble_q_function.py:108-134,177-184):
    (B, 1) minimum over members, or for QR members the (B, n_quantiles) vector of the member with the smallest mean."""
    th = quantiles_continuous(p, x, action)
    if th.shape[2] == 1:
        return reduce_ensemble(th, "min")
    return reduce_quantile_ensemble_min(th)


def q_discrete(p: Params, x, reduction="mean", n_quantiles: Optional[int] = None) -> torch.Tensor:
    """EnsembleDiscreteQFunction.forward (ensemble_q_function.py:139-146); QR members return the mean over their
    quantiles (DiscreteQRQFunction.forward, qr_q_function.py:44-48)."""
    vals = []
    for i in range(n_members(p)):
        pre = f"_q_funcs.{i}._encoder."
        h = pixel_forward(p, pre, x) if f"{pre}_convs.0.weight" in p else mlp_forward(p, pre, x)
        q = F.linear(h, p[f"_q_funcs.{i}._fc.weight"], p[f"_q_funcs.{i}._fc.bias"])
        if n_quantiles:
            q = q.view(x.shape[0], -1, n_quantiles).mean(dim=2)
        vals.append(q.view(1, x.shape[0], -1))
    return reduce_ensemble(torch.cat(vals, dim=0), reduction)


def quantiles_discrete(p: Params, x, n_quantiles: int) -> torch.Tensor:
    """DiscreteQRQFunction._compute_quantiles of every member: (E, B, A, n_quantiles) (qr_q_function.py:38-42)."""
    vals = []
    for i in range(n_members(p)):
        pre = f"_q_funcs.{i}._encoder."
        h = pixel_forward(p, pre, x) if f"{pre}_convs.0.weight" in p else mlp_forward(p, pre, x)
        q = F.linear(h, p[f"_q_funcs.{i}._fc.weight"], p[f"_q_funcs.{i}._fc.bias"])
        vals.append(q.view(1, x.shape[0], -1, n_quantiles))
    return torch.cat(vals, dim=0)


def make_taus(n_quantiles: int) -> torch.Tensor:
    """``_make_taus`` (qr_q_function.py:15-19): mid-points of the n uniform probability bins."""
    steps = torch.arange(n_quantiles, dtype=torch.float32)
    taus = ((steps + 1).float() / n_quantiles).view(1, -1)
    taus_dot = (steps.float() / n_quantiles).view(1, -1)
    return (taus + taus_dot) / 2.0


def quantile_huber_loss(quantiles, rew, target, term, taus, gamma) -> torch.Tensor:
    """compute_quantile_loss / compute_quantile_huber_loss (q_functions/utility.py:35-61): per sample
    mean_j sum_i |tau_i - 1[y_j - theta_i < 0]| * huber(y_j - theta_i)."""
    B, n = quantiles.shape
    y = rew + gamma * target * (1 - term)
    th, ey, et = quantiles.view(B, 1, -1), y.view(B, -1, 1), taus.view(-1, 1, n)
    hub = huber(th, ey)
    delta = ((ey - th).detach() < 0.0).float()
    return ((et - delta).abs() * hub).sum(dim=2).mean(dim=1)


def td_error_discrete_qr(p: Params, obs, act_long, rew, target, term, gamma, n_quantiles) -> torch.Tensor:
    """DiscreteQRQFunction.compute_error (qr_q_function.py:50-78) summed over members
    (EnsembleQFunction.compute_error, ensemble_q_function.py:81-106)."""
    assert target.shape == (obs.shape[0], n_quantiles)
    th = quantiles_discrete(p, obs, n_quantiles)
    one_hot = F.one_hot(act_long.view(-1), num_classes=th.shape[2]).view(-1, th.shape[2], 1).float()
    taus = make_taus(n_quantiles)
    total = torch.tensor(0.0)
    for i in range(th.shape[0]):
        picked = (th[i] * one_hot).sum(dim=1)  # pick_quantile_value_by_action (utility.py:17-24)
        total = total + quantile_huber_loss(picked, rew, target, term, taus, gamma).view(-1, 1).mean()
    return total


def reduce_quantile_ensemble_min(y: torch.Tensor) -> torch.Tensor:
    """``_reduce_quantile_ensemble(.., "min")`` for (E, B, n) (ensemble_q_function.py:27-52): per sample the
    quantiles of the member whose mean is smallest."""
    idx = y.mean(dim=-1).min(dim=0).indices
    return y.transpose(0, 1)[torch.arange(y.shape[1]), idx]


def td_error_continuous(p: Params, obs, act, rew, target, term, gamma) -> torch.Tensor:
    """EnsembleQFunction.compute_error (ensemble_q_function.py:81-106): sum over members of
    per-member batch-mean MSE (mean_q_function.py:74-87)."""
    assert target.ndim == 2
    total = torch.tensor(0.0)
    th = quantiles_continuous(p, obs, act)
    if th.shape[2] > 1:  # ContinuousQRQFunction.compute_error (qr_q_function.py:124-148)
        assert target.shape == (obs.shape[0], th.shape[2])
        taus = make_taus(th.shape[2])
        for i in range(th.shape[0]):
            total = total + quantile_huber_loss(th[i], rew, target, term, taus, gamma).view(-1, 1).mean()
        return total
    y = rew + gamma * target * (1 - term)
    for i in range(th.shape[0]):
        total = total + F.mse_loss(th[i], y, reduction="none").mean()
    return total


def huber(y, target, beta=1.0):
    """compute_huber_loss (q_functions/utility.py:27-32)."""
    diff = target - y
    cond = diff.detach().abs() < beta
    return torch.where(cond, 0.5 * diff ** 2, beta * (diff.abs() - 0.5 * beta))


def td_error_discrete(p: Params, obs, act_long, rew, target, term, gamma) -> torch.Tensor:
    """DiscreteMeanQFunction.compute_error (mean_q_function.py:26-42), summed over members."""
    q = q_discrete(p, obs, "none")
    one_hot = F.one_hot(act_long.view(-1), num_classes=q.shape[2]).float()
    y = rew + gamma * target * (1 - term)
    total = torch.tensor(0.0)
    for i in range(q.shape[0]):
        value = (q[i] * one_hot).sum(dim=1, keepdim=True)
        total = total + huber(value, y).mean()
    return total


LOG2 = math.log(2)


def squashed_log_prob(mu, std, raw):
    """SquashedGaussianDistribution._log_prob_from_raw_y (distributions.py:133-135)."""
    normal_lp = -((raw - mu) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))
    jacob = 2 * (LOG2 - raw - F.softplus(-2 * raw))
    return (normal_lp - jacob).sum(dim=-1, keepdim=True)


def policy_dist(p: Params, x, min_logstd=-20.0, max_logstd=2.0):
    """NormalPolicy.dist (policies.py:167-181): mu, std=exp(clamp(logstd))."""
    h = mlp_forward(p, "_encoder.", x)
    mu = F.linear(h, p["_mu.weight"], p["_mu.bias"])
    logstd = F.linear(h, p["_logstd.weight"], p["_logstd.bias"]).clamp(min_logstd, max_logstd)
    return mu, logstd.exp()


def policy_sample_with_log_prob(p: Params, x, eps):
    """sample_with_log_prob (policies.py:196-200; distributions.py:103-106); eps (B,A)."""
    mu, std = policy_dist(p, x)
    raw = mu + eps * std
    return torch.tanh(raw), squashed_log_prob(mu, std, raw)


def policy_sample_n_with_log_prob(p: Params, x, eps):
    """sample_n_with_log_prob (policies.py:202-213; distributions.py:116-121); eps (N,B,A)
    -> actions (B,N,A), log_probs (B,N,1)."""
    mu, std = policy_dist(p, x)
    raw = mu.unsqueeze(0) + eps * std.unsqueeze(0)
    lp = squashed_log_prob(mu.unsqueeze(0), std.unsqueeze(0), raw)
    return torch.tanh(raw).transpose(0, 1), lp.transpose(0, 1)


def policy_best_action(p: Params, x):
    """best_action -> mean_with_log_prob -> tanh(mu) (policies.py:247-249, distributions.py:126-127)."""
    mu, _ = policy_dist(p, x)
    return torch.tanh(mu)


def deterministic_policy(p: Params, x):
    """DeterministicPolicy.forward (policies.py:57-59)."""
    return torch.tanh(F.linear(mlp_forward(p, "_encoder.", x), p["_fc.weight"], p["_fc.bias"]))


def residual_policy(p: Params, x, action, scale):
    """DeterministicResidualPolicy.forward (policies.py:94-97)."""
    h = mlp_forward(p, "_encoder.", torch.cat([x, action], dim=1))
    res = scale * torch.tanh(F.linear(h, p["_fc.weight"], p["_fc.bias"]))
    return (action + res).clamp(-1.0, 1.0)


def vae_decode(p: Params, x, latent):
    """ConditionalVAE.decode (imitators.py:70-72)."""
    h = mlp_forward(p, "_decoder_encoder.", torch.cat([x, latent], dim=1))
    return torch.tanh(F.linear(h, p["_fc.weight"], p["_fc.bias"]))


def vae_error(p: Params, x, action, eps, beta, min_logstd=-4.0, max_logstd=15.0):
    """ConditionalVAE.compute_error (imitators.py:80-86); eps (B,latent)."""
    h = mlp_forward(p, "_encoder_encoder.", torch.cat([x, action], dim=1))
    mu = F.linear(h, p["_mu.weight"], p["_mu.bias"])
    logstd = F.linear(h, p["_logstd.weight"], p["_logstd.bias"]).clamp(min_logstd, max_logstd)
    std = logstd.exp()
    # kl_divergence(Normal(mu,std), Normal(0,1)) (torch/distributions/kl.py _kl_normal_normal)
    var_ratio = std.pow(2)
    t1 = mu.pow(2)
    kl = 0.5 * (var_ratio + t1 - 1 - var_ratio.log())
    y = vae_decode(p, x, mu + eps * std)
    return F.mse_loss(y, action) + beta * kl.mean()


# --------------------------------------------------------------------------- optimiser / sync
def make_adam(params: Params, lr: float, betas=(0.9, 0.999), eps=1e-8):
    """AdamFactory defaults (d3rlpy/models/optimizers.py:106-138)."""
    return torch.optim.Adam(list(params.values()), lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)


def soft_sync(targ: Params, src: Params, tau: float):
    """torch_utility.soft_sync (torch_utility.py:27-33): mul_(1-tau) then add_(tau*p)."""
    with torch.no_grad():
        for k in src:
            targ[k].mul_(1 - tau)
            targ[k].add_(tau * src[k])


def hard_sync(targ: Params, src: Params):
    """torch_utility.hard_sync (torch_utility.py:36-41)."""
    with torch.no_grad():
        for k in src:
            targ[k].copy_(src[k])


# --------------------------------------------------------------------------- batches / scalers
class Batch:
    """TorchMiniBatch (torch_utility.py:152-221): every field float32; scaler applied to
    observations/next_observations, action scaler to actions, reward scaler to rewards (oracle/scalers.py)."""

    def __init__(self, arrays: dict, scaler=None, reward_scaler=None, action_scaler=None):
        f = lambda a: torch.tensor(np.asarray(a)).float()
        self.observations = f(arrays["observations"])
        self.actions = f(arrays["actions"])
        self.rewards = f(arrays["rewards"])
        self.next_observations = f(arrays["next_observations"])
        self.terminals = f(arrays["terminals"])
        self.n_steps = f(arrays["n_steps"])
        if scaler is not None:
            self.observations = scaler(self.observations)
            self.next_observations = scaler(self.next_observations)
        if action_scaler is not None:
            self.actions = action_scaler(self.actions)
        if reward_scaler is not None:
            self.rewards = reward_scaler(self.rewards)


def standard_scaler(mean, std, eps=1e-3):
    """StandardScaler.transform (preprocessing/scalers.py:350-354)."""
    mean = torch.tensor(np.asarray(mean), dtype=torch.float32).reshape(1, -1)
    std = torch.tensor(np.asarray(std), dtype=torch.float32).reshape(1, -1)
    return lambda x: (x - mean) / (std + eps)


def pixel_scaler():
    """PixelScaler.transform (preprocessing/scalers.py:109-110)."""
    return lambda x: x.float() / 255.0


def clip_reward_scaler(low, high, multiplier=1.0):
    """ClipRewardScaler.transform (preprocessing/reward_scalers.py:176-177)."""
    return lambda r: multiplier * r.clamp(low, high)


# --------------------------------------------------------------------------- algorithms
class _Algo:
    grad_step = 0

    def update(self, batch: Batch, noise: Noise) -> Dict[str, float]:
        """LearnableBase.update (d3rlpy/base.py:746-758)."""
        m = self._update(batch, noise)
        self.grad_step += 1
        return m


class TD3PlusBC(_Algo):
    """TD3PlusBC._update (algos/td3_plus_bc.py:177-192) over TD3PlusBCImpl."""

    def __init__(self, obs, act, hidden=(256, 256), n_critics=2, actor_lr=3e-4, critic_lr=3e-4,
                 gamma=0.99, tau=0.005, sigma=0.2, clip=0.5, alpha=2.5, update_actor_interval=2,
                 seed=0, policy=None, critics=None):
        gen = torch.Generator().manual_seed(seed)
        self.q = clone_params(critics if critics is not None else make_critics(obs, act, hidden, n_critics, gen))
        self.pi = clone_params(policy if policy is not None else make_deterministic_policy(obs, act, hidden, gen))
        self.targ_q = clone_params(self.q, False)
        self.targ_pi = clone_params(self.pi, False)
        self.critic_optim = make_adam(self.q, critic_lr)
        self.actor_optim = make_adam(self.pi, actor_lr)
        self.gamma, self.tau, self.sigma, self.clip, self.alpha = gamma, tau, sigma, clip, alpha
        self.update_actor_interval = update_actor_interval
        self.grad_step = 0

    def compute_target(self, b: Batch, noise: Noise):
        """TD3Impl.compute_target (algos/torch/td3_impl.py:61-78)."""
        with torch.no_grad():
            action = deterministic_policy(self.targ_pi, b.next_observations)
            n = noise.normal(*action.shape)
            clipped_noise = (self.sigma * n).clamp(-self.clip, self.clip)
            a = (action + clipped_noise).clamp(-1.0, 1.0)
            return q_target_continuous(self.targ_q, b.next_observations, a)

    def compute_critic_loss(self, b: Batch, q_tpn):
        """DDPGBaseImpl.compute_critic_loss (algos/torch/ddpg_impl.py:154-165)."""
        return td_error_continuous(self.q, b.observations, b.actions, b.rewards, q_tpn, b.terminals,
                                   self.gamma ** b.n_steps)

    def compute_actor_loss(self, b: Batch):
        """TD3PlusBCImpl.compute_actor_loss (algos/torch/td3_plus_bc_impl.py:64-70)."""
        action = deterministic_policy(self.pi, b.observations)
        q_t = q_continuous(self.q, b.observations, action, "none")[0]
        lam = self.alpha / (q_t.abs().mean()).detach()
        return lam * -q_t.mean() + ((b.actions - action) ** 2).mean()

    def update_critic(self, b, noise):
        self.critic_optim.zero_grad()
        loss = self.compute_critic_loss(b, self.compute_target(b, noise))
        loss.backward()
        self.critic_optim.step()
        return float(loss.detach())

    def update_actor(self, b):
        self.actor_optim.zero_grad()
        loss = self.compute_actor_loss(b)
        loss.backward()
        for v in self.q.values():  # critic grads produced here are discarded by the next zero_grad
            v.grad = None
        self.actor_optim.step()
        return float(loss.detach())

    def _update(self, b, noise):
        m = {"critic_loss": self.update_critic(b, noise)}
        if self.grad_step % self.update_actor_interval == 0:
            m["actor_loss"] = self.update_actor(b)
            soft_sync(self.targ_q, self.q, self.tau)
            soft_sync(self.targ_pi, self.pi, self.tau)
        return m


class CQL(_Algo):
    """CQL._update (algos/cql.py:234-258) over CQLImpl/SACImpl."""

    def __init__(self, obs, act, hidden=(256, 256, 256), n_critics=2, actor_lr=1e-4, critic_lr=3e-4,
                 temp_lr=1e-4, alpha_lr=1e-4, gamma=0.99, tau=0.005, initial_temperature=1.0,
                 initial_alpha=1.0, alpha_threshold=10.0, conservative_weight=5.0, n_action_samples=10,
                 soft_q_backup=False, seed=0, policy=None, critics=None, actor_hidden=None):
        gen = torch.Generator().manual_seed(seed)
        self.q = clone_params(critics if critics is not None else make_critics(obs, act, hidden, n_critics, gen))
        self.pi = clone_params(policy if policy is not None
                               else make_squashed_normal_policy(obs, act, actor_hidden or hidden, gen))
        self.targ_q = clone_params(self.q, False)
        self.targ_pi = clone_params(self.pi, False)
        self.log_temp = {"_parameter": torch.full((1, 1), math.log(initial_temperature)).requires_grad_(True)}
        self.log_alpha = {"_parameter": torch.full((1, 1), math.log(initial_alpha)).requires_grad_(True)}
        self.critic_optim = make_adam(self.q, critic_lr)
        self.actor_optim = make_adam(self.pi, actor_lr)
        self.temp_optim = make_adam(self.log_temp, temp_lr)
        self.alpha_optim = make_adam(self.log_alpha, alpha_lr)
        self.temp_lr, self.alpha_lr = temp_lr, alpha_lr
        self.gamma, self.tau = gamma, tau
        self.alpha_threshold, self.conservative_weight = alpha_threshold, conservative_weight
        self.n, self.soft_q_backup, self.act = n_action_samples, soft_q_backup, act
        self.n_critics = n_members(self.q)
        self.grad_step = 0

    # -- SAC pieces (algos/torch/sac_impl.py:114-162)
    def update_temp(self, b, noise):
        self.temp_optim.zero_grad()
        with torch.no_grad():
            _, log_prob = policy_sample_with_log_prob(self.pi, b.observations, noise.normal(*b.actions.shape))
            targ_temp = log_prob - self.act
        loss = -(self.log_temp["_parameter"].exp() * targ_temp).mean()
        loss.backward()
        self.temp_optim.step()
        return float(loss.detach()), float(self.log_temp["_parameter"].exp().detach()[0][0])

    def compute_actor_loss(self, b, noise):
        action, log_prob = policy_sample_with_log_prob(self.pi, b.observations, noise.normal(*b.actions.shape))
        entropy = self.log_temp["_parameter"].exp() * log_prob
        q_t = q_continuous(self.q, b.observations, action, "min")
        return (entropy - q_t).mean()

    # -- CQL pieces (algos/torch/cql_impl.py:110-243)
    def _policy_is_values(self, policy_obs, value_obs, noise):
        B = value_obs.shape[0]
        with torch.no_grad():
            acts, lps = policy_sample_n_with_log_prob(self.pi, policy_obs, noise.normal(self.n, B, self.act))
        flat_obs = value_obs.expand(self.n, *value_obs.shape).transpose(0, 1).reshape(-1, value_obs.shape[1])
        vals = q_continuous(self.q, flat_obs, acts.reshape(-1, self.act), "none").view(self.n_critics, B, self.n)
        return vals - lps.reshape(1, -1, self.n)

    def _random_is_values(self, obs, noise):
        B = obs.shape[0]
        flat_obs = obs.expand(self.n, *obs.shape).transpose(0, 1).reshape(-1, obs.shape[1])
        rand = noise.uniform(B * self.n, self.act)
        vals = q_continuous(self.q, flat_obs, rand, "none").view(self.n_critics, B, self.n)
        return vals - math.log(0.5 ** self.act)

    def conservative_loss(self, obs_t, act_t, obs_tp1, noise):
        v_t = self._policy_is_values(obs_t, obs_t, noise)
        v_tp1 = self._policy_is_values(obs_tp1, obs_t, noise)
        v_r = self._random_is_values(obs_t, noise)
        target_values = torch.cat([v_t, v_tp1, v_r], dim=2)
        lse = torch.logsumexp(target_values, dim=2, keepdim=True)
        data_values = q_continuous(self.q, obs_t, act_t, "none")
        loss = lse.mean(dim=0).mean() - data_values.mean(dim=0).mean()
        scaled = self.conservative_weight * loss
        clipped_alpha = self.log_alpha["_parameter"].exp().clamp(0, 1e6)[0][0]
        return clipped_alpha * (scaled - self.alpha_threshold)

    def update_alpha(self, b, noise):
        self.alpha_optim.zero_grad()
        loss = -self.conservative_loss(b.observations, b.actions, b.next_observations, noise)
        loss.backward()
        for v in self.q.values():
            v.grad = None
        self.alpha_optim.step()
        return float(loss.detach()), float(self.log_alpha["_parameter"].exp().detach()[0][0])

    def compute_target(self, b, noise):
        with torch.no_grad():
            if self.soft_q_backup:
                a, lp = policy_sample_with_log_prob(self.pi, b.next_observations, noise.normal(*b.actions.shape))
                ent = self.log_temp["_parameter"].exp() * lp
                return q_continuous(self.targ_q, b.next_observations, a, "min") - ent
            a = policy_best_action(self.pi, b.next_observations)
            return q_continuous(self.targ_q, b.next_observations, a, "min")

    def compute_critic_loss(self, b, q_tpn, noise):
        td = td_error_continuous(self.q, b.observations, b.actions, b.rewards, q_tpn, b.terminals,
                                 self.gamma ** b.n_steps)
        return td + self.conservative_loss(b.observations, b.actions, b.next_observations, noise)

    def update_critic(self, b, noise):
        self.critic_optim.zero_grad()
        q_tpn = self.compute_target(b, noise)
        loss = self.compute_critic_loss(b, q_tpn, noise)
        loss.backward()
        self.log_alpha["_parameter"].grad = None
        self.critic_optim.step()
        return float(loss.detach())

    def update_actor(self, b, noise):
        self.actor_optim.zero_grad()
        loss = self.compute_actor_loss(b, noise)
        loss.backward()
        for v in self.q.values():
            v.grad = None
        self.log_temp["_parameter"].grad = None
        self.actor_optim.step()
        return float(loss.detach())

    def _update(self, b, noise):
        m = {}
        if self.temp_lr > 0:
            m["temp_loss"], m["temp"] = self.update_temp(b, noise)
        if self.alpha_lr > 0:
            m["alpha_loss"], m["alpha"] = self.update_alpha(b, noise)
        m["critic_loss"] = self.update_critic(b, noise)
        m["actor_loss"] = self.update_actor(b, noise)
        soft_sync(self.targ_q, self.q, self.tau)
        soft_sync(self.targ_pi, self.pi, self.tau)
        return m


class SAC(CQL):
    """SAC._update (algos/sac.py:177-198) over SACImpl (algos/torch/sac_impl.py:88-162): the update CQL extends --
    temperature step, TD critic loss against the soft target min_e Q'(s', a') - exp(log_temp) log pi(a'|s') with a
    sampled a', actor loss, both soft syncs.  No conservative term, no alpha."""

    def __init__(self, obs, act, hidden=(256, 256), n_critics=2, actor_lr=3e-4, critic_lr=3e-4, temp_lr=3e-4,
                 gamma=0.99, tau=0.005, initial_temperature=1.0, seed=0, policy=None, critics=None):
        super().__init__(obs, act, hidden=hidden, n_critics=n_critics, actor_lr=actor_lr, critic_lr=critic_lr,
                         temp_lr=temp_lr, alpha_lr=0.0, gamma=gamma, tau=tau,
                         initial_temperature=initial_temperature, n_action_samples=0, soft_q_backup=True, seed=seed,
                         policy=policy, critics=critics)

    def compute_critic_loss(self, b, q_tpn, noise):
        """DDPGBaseImpl.compute_critic_loss (algos/torch/ddpg_impl.py:154-165)."""
        return td_error_continuous(self.q, b.observations, b.actions, b.rewards, q_tpn, b.terminals,
                                   self.gamma ** b.n_steps)


class TD3(TD3PlusBC):
    """TD3._update (algos/td3.py:161-176) over TD3Impl (algos/torch/td3_impl.py): TD3+BC without the behaviour-cloning
    term -- actor loss -Q_0(s, pi(s)).mean() (DDPGImpl.compute_actor_loss, algos/torch/ddpg_impl.py:268-273)."""

    def __init__(self, obs, act, hidden=(256, 256), n_critics=2, actor_lr=3e-4, critic_lr=3e-4, gamma=0.99,
                 tau=0.005, sigma=0.2, clip=0.5, update_actor_interval=2, seed=0, policy=None, critics=None):
        super().__init__(obs, act, hidden=hidden, n_critics=n_critics, actor_lr=actor_lr, critic_lr=critic_lr,
                         gamma=gamma, tau=tau, sigma=sigma, clip=clip, alpha=0.0,
                         update_actor_interval=update_actor_interval, seed=seed, policy=policy, critics=critics)

    def compute_actor_loss(self, b: Batch):
        action = deterministic_policy(self.pi, b.observations)
        return -q_continuous(self.q, b.observations, action, "none")[0].mean()


class DDPG(TD3):
    """DDPG._update (algos/ddpg.py:168-176) over DDPGImpl (algos/torch/ddpg_impl.py:255-288): no target smoothing (and
    no noise draw), actor and both soft syncs on every step."""

    def __init__(self, obs, act, hidden=(256, 256), n_critics=1, actor_lr=3e-4, critic_lr=3e-4, gamma=0.99,
                 tau=0.005, seed=0, policy=None, critics=None):
        super().__init__(obs, act, hidden=hidden, n_critics=n_critics, actor_lr=actor_lr, critic_lr=critic_lr,
                         gamma=gamma, tau=tau, sigma=0.0, clip=0.5, update_actor_interval=1, seed=seed, policy=policy,
                         critics=critics)

    def compute_target(self, b: Batch, noise: Noise):
        """DDPGImpl.compute_target (algos/torch/ddpg_impl.py:275-284)."""
        with torch.no_grad():
            action = deterministic_policy(self.targ_pi, b.next_observations)
            return q_target_continuous(self.targ_q, b.next_observations, action.clamp(-1.0, 1.0))


def make_non_squashed_normal_policy(obs: int, act: int, hidden, gen) -> Params:
    """NonSquashedNormalPolicy with ``use_std_parameter=True`` (policies.py:127-158,274-290): the ``_logstd``
    nn.Parameter (zeros, shape (1, A)) is registered on the module itself, so it precedes the sub-modules in
    ``state_dict()`` / ``parameters()`` order."""
    p: Params = OrderedDict([("_logstd", torch.zeros(1, act))])
    p.update(make_mlp("_encoder.", obs, hidden, gen))
    p.update(make_head("_mu", act, hidden[-1], gen))
    return p


def make_value_function(obs: int, hidden, gen) -> Params:
    """ValueFunction (v_functions.py:10-21): VectorEncoder + Linear(feature, 1)."""
    p = make_mlp("_encoder.", obs, hidden, gen)
    p.update(make_head("_fc", 1, hidden[-1], gen))
    return p


def value_function(p: Params, x):
    return F.linear(mlp_forward(p, "_encoder.", x), p["_fc.weight"], p["_fc.bias"])


def non_squashed_policy_dist(p: Params, x, min_logstd=-5.0, max_logstd=2.0):
    """NormalPolicy.dist with ``squash_distribution=False`` (policies.py:168-181): Normal(tanh(mu), exp(logstd)) with
    logstd = min + sigmoid(_logstd) * (max - min) (``get_logstd_parameter``, policies.py:248-253)."""
    mu = F.linear(mlp_forward(p, "_encoder.", x), p["_mu.weight"], p["_mu.bias"])
    logstd = min_logstd + torch.sigmoid(p["_logstd"]) * (max_logstd - min_logstd)
    return torch.distributions.Normal(torch.tanh(mu), logstd.exp())


class IQL(_Algo):
    """IQL._update (algos/iql.py:186-199) over IQLImpl (algos/torch/iql_impl.py:74-200): expectile value regression,
    TD on V(s'), advantage-weighted Gaussian log-likelihood; ONE Adam over the critics and the value function."""

    def __init__(self, obs, act, hidden=(256, 256), n_critics=2, actor_lr=3e-4, critic_lr=3e-4, gamma=0.99, tau=0.005,
                 expectile=0.7, weight_temp=3.0, max_weight=100.0, seed=0, policy=None, critics=None, value=None):
        gen = torch.Generator().manual_seed(seed)
        self.q = clone_params(critics if critics is not None else make_critics(obs, act, hidden, n_critics, gen))
        self.pi = clone_params(policy if policy is not None else make_non_squashed_normal_policy(obs, act, hidden, gen))
        self.v = clone_params(value if value is not None else make_value_function(obs, hidden, gen))
        self.targ_q, self.targ_pi = clone_params(self.q, False), clone_params(self.pi, False)
        self.critic_optim = torch.optim.Adam(list(self.q.values()) + list(self.v.values()), lr=critic_lr)
        self.actor_optim = make_adam(self.pi, actor_lr)
        self.gamma, self.tau = gamma, tau
        self.expectile, self.weight_temp, self.max_weight = expectile, weight_temp, max_weight
        self.grad_step = 0

    def compute_target(self, b: Batch):
        with torch.no_grad():
            return value_function(self.v, b.next_observations)

    def compute_critic_loss(self, b: Batch, q_tpn):
        return td_error_continuous(self.q, b.observations, b.actions, b.rewards, q_tpn, b.terminals,
                                   self.gamma ** b.n_steps)

    def compute_value_loss(self, b: Batch):
        q_t = q_continuous(self.targ_q, b.observations, b.actions, "min")
        v_t = value_function(self.v, b.observations)
        diff = q_t.detach() - v_t
        weight = (self.expectile - (diff < 0.0).float()).abs().detach()
        return (weight * (diff ** 2)).mean()

    def compute_weight(self, b: Batch):
        q_t = q_continuous(self.targ_q, b.observations, b.actions, "min")
        v_t = value_function(self.v, b.observations)
        return (self.weight_temp * (q_t - v_t)).exp().clamp(max=self.max_weight)

    def compute_actor_loss(self, b: Batch):
        log_probs = non_squashed_policy_dist(self.pi, b.observations).log_prob(b.actions).sum(dim=-1, keepdim=True)
        with torch.no_grad():
            weight = self.compute_weight(b)
        return -(weight * log_probs).mean()

    def _update(self, b, noise=None):
        self.critic_optim.zero_grad()
        q_loss = self.compute_critic_loss(b, self.compute_target(b))
        v_loss = self.compute_value_loss(b)
        (q_loss + v_loss).backward()
        self.critic_optim.step()
        self.actor_optim.zero_grad()
        a_loss = self.compute_actor_loss(b)
        a_loss.backward()
        self.actor_optim.step()
        soft_sync(self.targ_q, self.q, self.tau)
        return {"critic_loss": float(q_loss.detach()), "value_loss": float(v_loss.detach()),
                "actor_loss": float(a_loss.detach())}


class AWAC(_Algo):
    """AWAC._update (algos/awac.py:176-191) over AWACImpl (algos/torch/awac_impl.py:18-154), which is SACImpl with a
    frozen temperature exp(log 1e-20), a NonSquashedNormalPolicy whose logstd is a parameter squashed into [-6, 0], and
    an actor Adam with weight_decay 1e-4 (algos/awac.py:105).  Oracle only: the CUDA path for AWAC is not built yet
    (DESIGN.md section 6b); this class and its golden case pin what that path has to reproduce."""

    MIN_LOGSTD, MAX_LOGSTD = -6.0, 0.0

    def __init__(self, obs, act, hidden=(256, 256), n_critics=2, actor_lr=3e-4, critic_lr=3e-4, gamma=0.99, tau=0.005,
                 lam=1.0, n_action_samples=1, update_actor_interval=1, actor_weight_decay=1e-4, seed=0, policy=None,
                 critics=None):
        gen = torch.Generator().manual_seed(seed)
        self.q = clone_params(critics if critics is not None else make_critics(obs, act, hidden, n_critics, gen))
        self.pi = clone_params(policy if policy is not None else make_non_squashed_normal_policy(obs, act, hidden, gen))
        self.targ_q, self.targ_pi = clone_params(self.q, False), clone_params(self.pi, False)
        self.critic_optim = make_adam(self.q, critic_lr)
        self.actor_optim = torch.optim.Adam(list(self.pi.values()), lr=actor_lr, weight_decay=actor_weight_decay)
        self.log_temp = torch.full((1, 1), math.log(1e-20))   # temp_learning_rate = 0 and never stepped
        self.gamma, self.tau, self.lam, self.n, self.act = gamma, tau, lam, n_action_samples, act
        self.update_actor_interval = update_actor_interval
        self.grad_step = 0

    def _dist(self, x):
        return non_squashed_policy_dist(self.pi, x, self.MIN_LOGSTD, self.MAX_LOGSTD)

    def compute_target(self, b: Batch, noise: Noise):
        """SACImpl.compute_target (sac_impl.py:148-162) with GaussianDistribution.sample_with_log_prob
        (distributions.py:52-57,83-84): a' = clamp(tanh(mu) + std * eps, -1, 1)."""
        with torch.no_grad():
            dist = self._dist(b.next_observations)
            action = (dist.loc + noise.normal(*b.actions.shape) * dist.scale).clamp(-1.0, 1.0)
            log_prob = dist.log_prob(action).sum(dim=-1, keepdim=True)
            entropy = self.log_temp.exp() * log_prob
            return q_target_continuous(self.targ_q, b.next_observations, action) - entropy

    def compute_critic_loss(self, b: Batch, q_tpn):
        return td_error_continuous(self.q, b.observations, b.actions, b.rewards, q_tpn, b.terminals,
                                   self.gamma ** b.n_steps)

    def compute_weights(self, b: Batch, noise: Noise):
        """AWACImpl._compute_weights (awac_impl.py:118-154): softmax over the batch of (Q - V) / lam, times the batch
        size; V(s) = mean over n sampled actions of min_e Q_e(s, a_n)."""
        with torch.no_grad():
            B = b.observations.shape[0]
            q_values = q_continuous(self.q, b.observations, b.actions, "min")
            dist = self._dist(b.observations)
            acts_T = (dist.loc + noise.normal(self.n, B, self.act) * dist.scale).clamp(-1.0, 1.0)   # (n, B, A)
            flat_actions = acts_T.transpose(0, 1).reshape(-1, self.act)
            flat_obs = b.observations.view(B, 1, -1).expand(B, self.n, b.observations.shape[1]).reshape(B * self.n, -1)
            v_values = q_continuous(self.q, flat_obs, flat_actions, "min").view(B, -1, 1).mean(dim=1)
            adv = (q_values - v_values).view(-1)
            return F.softmax(adv / self.lam, dim=0).view(-1, 1) * adv.numel()

    def compute_actor_loss(self, b: Batch, noise: Noise):
        log_probs = self._dist(b.observations).log_prob(b.actions).sum(dim=-1, keepdim=True)
        return -(log_probs * self.compute_weights(b, noise)).sum()

    def _update(self, b, noise):
        self.critic_optim.zero_grad()
        loss = self.compute_critic_loss(b, self.compute_target(b, noise))
        loss.backward()
        self.critic_optim.step()
        m = {"critic_loss": float(loss.detach())}
        if self.grad_step % self.update_actor_interval == 0:
            self.actor_optim.zero_grad()
            a_loss = self.compute_actor_loss(b, noise)
            a_loss.backward()
            self.actor_optim.step()
            logstd = self.MIN_LOGSTD + torch.sigmoid(self.pi["_logstd"]) * (self.MAX_LOGSTD - self.MIN_LOGSTD)
            m["actor_loss"], m["mean_std"] = float(a_loss.detach()), float(logstd.exp().mean().detach())
            soft_sync(self.targ_q, self.q, self.tau)
            soft_sync(self.targ_pi, self.pi, self.tau)
        return m


class CRR(_Algo):
    """CRR._update (algos/crr.py:226-244) over CRRImpl (algos/torch/crr_impl.py:17-191): DDPG-style critic step against
    a target action sampled from the TARGET policy, advantage-weighted Gaussian log-likelihood actor step (binary or
    clipped-exponential weights; state value = mean or max of Q over n sampled actions), hard or soft target updates.
    The policy is NonSquashedNormalPolicy with a logstd HEAD clamped to [-20, 2] (builders default), i.e. the same
    parameter layout as the squashed policy.  Oracle only: the CUDA path for CRR is not built yet."""

    def __init__(self, obs, act, hidden=(256, 256), n_critics=1, actor_lr=3e-4, critic_lr=3e-4, gamma=0.99, beta=1.0,
                 n_action_samples=4, advantage_type="mean", weight_type="exp", max_weight=20.0,
                 target_update_type="hard", tau=5e-3, target_update_interval=100, seed=0, policy=None, critics=None):
        gen = torch.Generator().manual_seed(seed)
        self.q = clone_params(critics if critics is not None else make_critics(obs, act, hidden, n_critics, gen))
        self.pi = clone_params(policy if policy is not None else make_squashed_normal_policy(obs, act, hidden, gen))
        self.targ_q, self.targ_pi = clone_params(self.q, False), clone_params(self.pi, False)
        self.critic_optim, self.actor_optim = make_adam(self.q, critic_lr), make_adam(self.pi, actor_lr)
        self.gamma, self.beta, self.n, self.act = gamma, beta, n_action_samples, act
        self.advantage_type, self.weight_type, self.max_weight = advantage_type, weight_type, max_weight
        self.target_update_type, self.tau, self.target_update_interval = target_update_type, tau, target_update_interval
        self.grad_step = 0

    @staticmethod
    def _dist(p: Params, x):
        """NormalPolicy.dist with squash_distribution=False and a logstd head (policies.py:160-181)."""
        h = mlp_forward(p, "_encoder.", x)
        mu = F.linear(h, p["_mu.weight"], p["_mu.bias"])
        logstd = F.linear(h, p["_logstd.weight"], p["_logstd.bias"]).clamp(-20.0, 2.0)
        return torch.distributions.Normal(torch.tanh(mu), logstd.exp())

    def compute_target(self, b: Batch, noise: Noise):
        """crr_impl.py:143-153: a' = targ_policy.sample(s') = clamp(tanh(mu) + std * eps, -1, 1) (distributions.py:52-53)."""
        with torch.no_grad():
            dist = self._dist(self.targ_pi, b.next_observations)
            action = (dist.loc + noise.normal(*b.actions.shape) * dist.scale).clamp(-1.0, 1.0)
            return q_target_continuous(self.targ_q, b.next_observations, action.clamp(-1.0, 1.0))

    def compute_advantage(self, b: Batch, noise: Noise):
        """crr_impl.py:104-141; `self._q_func(x, a)` reduces over members with the default "mean"."""
        with torch.no_grad():
            B = b.observations.shape[0]
            dist = self._dist(self.pi, b.observations)
            acts_T = (dist.loc + noise.normal(self.n, B, self.act) * dist.scale).clamp(-1.0, 1.0)
            flat_actions = acts_T.transpose(0, 1).reshape(-1, self.act)
            flat_obs = b.observations.view(B, 1, -1).expand(B, self.n, b.observations.shape[1]).reshape(B * self.n, -1)
            reshaped = q_continuous(self.q, flat_obs, flat_actions, "mean").view(B, -1, 1)
            values = reshaped.mean(dim=1) if self.advantage_type == "mean" else reshaped.max(dim=1).values
            return q_continuous(self.q, b.observations, b.actions, "mean") - values

    def compute_weight(self, b: Batch, noise: Noise):
        adv = self.compute_advantage(b, noise)
        if self.weight_type == "binary":
            return (adv > 0.0).float()
        return (adv / self.beta).exp().clamp(0.0, self.max_weight)

    def compute_actor_loss(self, b: Batch, noise: Noise):
        log_probs = self._dist(self.pi, b.observations).log_prob(b.actions).sum(dim=-1, keepdim=True)
        return -(log_probs * self.compute_weight(b, noise)).mean()

    def _update(self, b, noise):
        self.critic_optim.zero_grad()
        c_loss = td_error_continuous(self.q, b.observations, b.actions, b.rewards, self.compute_target(b, noise),
                                     b.terminals, self.gamma ** b.n_steps)
        c_loss.backward()
        self.critic_optim.step()
        self.actor_optim.zero_grad()
        a_loss = self.compute_actor_loss(b, noise)
        a_loss.backward()
        self.actor_optim.step()
        if self.target_update_type == "hard":
            if self.grad_step % self.target_update_interval == 0:
                hard_sync(self.targ_q, self.q)
                hard_sync(self.targ_pi, self.pi)
        else:
            soft_sync(self.targ_q, self.q, self.tau)
            soft_sync(self.targ_pi, self.pi, self.tau)
        return {"critic_loss": float(c_loss.detach()), "actor_loss": float(a_loss.detach())}


class PLAS(_Algo):
    """PLAS._update (algos/plas.py:189-206) over PLASImpl (algos/torch/plas_impl.py:25-168): a conditional VAE trained
    alone for `warmup_steps`, then TD3-style steps whose actions are decode(s, 2 * pi(s)) with a deterministic policy
    over the 2A-dimensional latent; the target mixes min / max over members with `lam`.  Oracle only."""

    def __init__(self, obs, act, hidden=(256, 256), vae_hidden=(256, 256), n_critics=2, actor_lr=1e-4, critic_lr=1e-3,
                 imitator_lr=1e-4, gamma=0.99, tau=0.005, lam=0.75, beta=0.5, update_actor_interval=1,
                 warmup_steps=500000, seed=0, policy=None, critics=None, imitator=None):
        gen = torch.Generator().manual_seed(seed)
        self.imitator = clone_params(imitator if imitator is not None else make_cvae(obs, act, 2 * act, vae_hidden, gen))
        self.q = clone_params(critics if critics is not None else make_critics(obs, act, hidden, n_critics, gen))
        self.pi = clone_params(policy if policy is not None else make_deterministic_policy(obs, 2 * act, hidden, gen))
        self.targ_q, self.targ_pi = clone_params(self.q, False), clone_params(self.pi, False)
        self.critic_optim, self.actor_optim = make_adam(self.q, critic_lr), make_adam(self.pi, actor_lr)
        self.imitator_optim = make_adam(self.imitator, imitator_lr)
        self.gamma, self.tau, self.lam, self.beta, self.act = gamma, tau, lam, beta, act
        self.update_actor_interval, self.warmup_steps = update_actor_interval, warmup_steps
        self.grad_step = 0

    def compute_target(self, b: Batch):
        with torch.no_grad():
            actions = vae_decode(self.imitator, b.next_observations, 2.0 * deterministic_policy(self.targ_pi, b.next_observations))
            return q_continuous(self.targ_q, b.next_observations, actions, "mix", self.lam)

    def compute_actor_loss(self, b: Batch):
        actions = vae_decode(self.imitator, b.observations, 2.0 * deterministic_policy(self.pi, b.observations))
        return -q_continuous(self.q, b.observations, actions, "none")[0].mean()

    def _update(self, b, noise):
        m = {}
        if self.grad_step < self.warmup_steps:
            self.imitator_optim.zero_grad()
            loss = vae_error(self.imitator, b.observations, b.actions,
                             noise.normal(b.observations.shape[0], 2 * self.act), self.beta)
            loss.backward()
            self.imitator_optim.step()
            m["imitator_loss"] = float(loss.detach())
            return m
        self.critic_optim.zero_grad()
        c_loss = td_error_continuous(self.q, b.observations, b.actions, b.rewards, self.compute_target(b), b.terminals,
                                     self.gamma ** b.n_steps)
        c_loss.backward()
        self.critic_optim.step()
        m["critic_loss"] = float(c_loss.detach())
        if self.grad_step % self.update_actor_interval == 0:
            self.actor_optim.zero_grad()
            a_loss = self.compute_actor_loss(b)
            a_loss.backward()
            for v in list(self.q.values()) + list(self.imitator.values()):
                v.grad = None
            self.actor_optim.step()
            m["actor_loss"] = float(a_loss.detach())
            soft_sync(self.targ_pi, self.pi, self.tau)
            soft_sync(self.targ_q, self.q, self.tau)
        return m


def max_with_n_actions_and_indices(targ_q: Params, x, actions, lam):
    """compute_max_with_n_actions_and_indices (q_functions/__init__.py:8-63): x (B, O), actions (B, N, A) -> per sample
    the lam-mix of the members' min / max value at the action maximising that mix, and the action's index."""
    B, N = actions.shape[0], actions.shape[1]
    flat_x = x.expand(N, *x.shape).transpose(0, 1).reshape(-1, x.shape[1])
    vals = q_continuous(targ_q, flat_x, actions.reshape(B * N, -1), "none")   # (E, B*N, 1)
    E = vals.shape[0]
    values = vals.view(E, B, N, -1).transpose(0, 1)                            # (B, E, N, 1)
    mean_values = values.mean(dim=3)
    max_values, max_idx = mean_values.max(dim=1)
    min_values, min_idx = mean_values.min(dim=1)
    action_idx = ((1.0 - lam) * max_values + lam * min_values).argmax(dim=1)
    flat_values = values.transpose(1, 2).reshape(B * N, E, -1)
    bn = torch.arange(B * N)
    mx = flat_values[bn, max_idx.reshape(-1)].view(B, N, -1)
    mn = flat_values[bn, min_idx.reshape(-1)].view(B, N, -1)
    return ((1.0 - lam) * mx + lam * mn)[torch.arange(B), action_idx], action_idx


class BEAR(CQL):
    """BEAR._update (algos/bear.py:279-309) over BEARImpl (algos/torch/bear_impl.py:41-329), which extends SACImpl:
    VAE step, temperature step, Lagrange step on the MMD constraint (log_alpha clamped to [-5, 10]), TD critic step
    against the best of n target samples (lam-mix over members) minus the entropy term of that sample, then an actor
    step on the MMD loss alone during warm-up and on SAC's loss + MMD loss afterwards.  MMD between n raw (pre-tanh)
    policy samples and n raw decoder samples with a Laplacian or Gaussian kernel.  Oracle only.
    Inherits the SAC pieces (update_temp, SAC actor loss) from the CQL oracle."""

    def __init__(self, obs, act, hidden=(256, 256), vae_hidden=(256, 256), n_critics=2, actor_lr=1e-4, critic_lr=3e-4,
                 imitator_lr=3e-4, temp_lr=1e-4, alpha_lr=1e-3, gamma=0.99, tau=0.005, initial_temperature=1.0,
                 initial_alpha=1.0, alpha_threshold=0.05, lam=0.75, n_target_samples=10, n_mmd_action_samples=4,
                 mmd_kernel="laplacian", mmd_sigma=20.0, vae_kl_weight=0.5, warmup_steps=40000, seed=0, policy=None,
                 critics=None, imitator=None):
        super().__init__(obs, act, hidden=hidden, n_critics=n_critics, actor_lr=actor_lr, critic_lr=critic_lr,
                         temp_lr=temp_lr, alpha_lr=alpha_lr, gamma=gamma, tau=tau,
                         initial_temperature=initial_temperature, initial_alpha=initial_alpha,
                         alpha_threshold=alpha_threshold, n_action_samples=0, seed=seed, policy=policy, critics=critics)
        gen = torch.Generator().manual_seed(seed + 1)
        self.imitator = clone_params(imitator if imitator is not None else make_cvae(obs, act, 2 * act, vae_hidden, gen))
        self.imitator_optim = make_adam(self.imitator, imitator_lr)
        self.lam, self.n_target, self.n_mmd = lam, n_target_samples, n_mmd_action_samples
        self.mmd_kernel, self.mmd_sigma, self.beta, self.warmup_steps = mmd_kernel, mmd_sigma, vae_kl_weight, warmup_steps

    def compute_mmd(self, obs, noise: Noise):
        """bear_impl.py:233-281; draw order: decoder latents (n*B, 2A), then policy eps (n, B, A)."""
        B, n, A = obs.shape[0], self.n_mmd, self.act
        with torch.no_grad():
            latent = noise.normal(n * B, 2 * A).clamp(-0.5, 0.5)
            flat_x = obs.expand(n, *obs.shape).reshape(-1, obs.shape[1])
            h = mlp_forward(self.imitator, "_decoder_encoder.", torch.cat([flat_x, latent], dim=1))
            behavior = F.linear(h, self.imitator["_fc.weight"], self.imitator["_fc.bias"]).view(n, B, -1).transpose(0, 1)
        mu, std = policy_dist(self.pi, obs)
        policy = (mu.unsqueeze(0) + noise.normal(n, B, A) * std.unsqueeze(0)).transpose(0, 1)
        if self.mmd_kernel == "gaussian":
            kernel = lambda x, y: (-((x - y) ** 2).sum(dim=3) / (2 * self.mmd_sigma)).exp()   # noqa: E731
        else:
            kernel = lambda x, y: (-(x - y).abs().sum(dim=3) / (2 * self.mmd_sigma)).exp()    # noqa: E731
        b1, p1 = behavior.reshape(B, -1, 1, A), policy.reshape(B, -1, 1, A)
        bT, pT = behavior.reshape(B, 1, -1, A), policy.reshape(B, 1, -1, A)
        mmd = kernel(p1, pT).mean(dim=[1, 2])
        mmd = mmd + kernel(b1, bT).mean(dim=[1, 2])
        mmd = mmd - 2 * kernel(p1, bT).mean(dim=[1, 2])
        return (mmd + 1e-6).sqrt().view(-1, 1)

    def compute_mmd_loss(self, obs, noise: Noise):
        alpha = self.log_alpha["_parameter"].exp()
        return (alpha * (self.compute_mmd(obs, noise) - self.alpha_threshold)).mean()

    def update_imitator(self, b, noise):
        self.imitator_optim.zero_grad()
        loss = vae_error(self.imitator, b.observations, b.actions, noise.normal(b.observations.shape[0], 2 * self.act),
                         self.beta)
        loss.backward()
        self.imitator_optim.step()
        return float(loss.detach())

    def update_alpha(self, b, noise):
        loss = -self.compute_mmd_loss(b.observations, noise)
        self.alpha_optim.zero_grad()
        loss.backward()
        for v in self.pi.values():
            v.grad = None
        self.alpha_optim.step()
        self.log_alpha["_parameter"].data.clamp_(-5.0, 10.0)
        return float(loss.detach()), float(self.log_alpha["_parameter"].exp().detach()[0][0])

    def compute_target(self, b, noise):
        with torch.no_grad():
            B = b.observations.shape[0]
            actions, log_probs = policy_sample_n_with_log_prob(self.pi, b.next_observations,
                                                               noise.normal(self.n_target, B, self.act))
            values, idx = max_with_n_actions_and_indices(self.targ_q, b.next_observations, actions, self.lam)
            return values - self.log_temp["_parameter"].exp() * log_probs[torch.arange(B), idx]

    def compute_critic_loss(self, b, q_tpn, noise):
        return td_error_continuous(self.q, b.observations, b.actions, b.rewards, q_tpn, b.terminals,
                                   self.gamma ** b.n_steps)

    def update_actor(self, b, noise):
        self.actor_optim.zero_grad()
        if self.grad_step < self.warmup_steps:
            loss = self.compute_mmd_loss(b.observations, noise)
        else:
            loss = self.compute_actor_loss(b, noise) + self.compute_mmd_loss(b.observations, noise)
        loss.backward()
        for v in list(self.q.values()) + [self.log_temp["_parameter"], self.log_alpha["_parameter"]]:
            v.grad = None
        self.actor_optim.step()
        return float(loss.detach())

    def _update(self, b, noise):
        m = {"imitator_loss": self.update_imitator(b, noise)}
        if self.temp_lr > 0:
            m["temp_loss"], m["temp"] = self.update_temp(b, noise)
        if self.alpha_lr > 0:
            m["alpha_loss"], m["alpha"] = self.update_alpha(b, noise)
        m["critic_loss"] = self.update_critic(b, noise)
        m["actor_loss"] = self.update_actor(b, noise)
        soft_sync(self.targ_pi, self.pi, self.tau)
        soft_sync(self.targ_q, self.q, self.tau)
        return m


class BCQ(_Algo):
    """BCQ._update (algos/bcq.py:261-279) over BCQImpl (algos/torch/bcq_impl.py:132-226)."""

    def __init__(self, obs, act, hidden=(400, 300), vae_hidden=(750, 750), n_critics=2, actor_lr=1e-3,
                 critic_lr=1e-3, imitator_lr=1e-3, gamma=0.99, tau=0.005, lam=0.75, n_action_samples=100,
                 action_flexibility=0.05, beta=0.5, update_actor_interval=1, rl_start_step=0, seed=0,
                 policy=None, critics=None, imitator=None):
        gen = torch.Generator().manual_seed(seed)
        self.imitator = clone_params(imitator if imitator is not None else make_cvae(obs, act, 2 * act, vae_hidden, gen))
        self.q = clone_params(critics if critics is not None else make_critics(obs, act, hidden, n_critics, gen))
        self.pi = clone_params(policy if policy is not None else make_residual_policy(obs, act, hidden, gen))
        self.targ_q = clone_params(self.q, False)
        self.targ_pi = clone_params(self.pi, False)
        self.critic_optim = make_adam(self.q, critic_lr)
        self.actor_optim = make_adam(self.pi, actor_lr)
        self.imitator_optim = make_adam(self.imitator, imitator_lr)
        self.gamma, self.tau, self.lam, self.n, self.flex, self.beta = gamma, tau, lam, n_action_samples, action_flexibility, beta
        self.update_actor_interval, self.rl_start_step, self.act = update_actor_interval, rl_start_step, act
        self.grad_step = 0

    def update_imitator(self, b, noise):
        self.imitator_optim.zero_grad()
        loss = vae_error(self.imitator, b.observations, b.actions, noise.normal(b.observations.shape[0], 2 * self.act), self.beta)
        loss.backward()
        self.imitator_optim.step()
        return float(loss.detach())

    def compute_target(self, b, noise):
        """bcq_impl.py:163-187,215-226 + compute_max_with_n_actions (q_functions/__init__.py:8-63)."""
        with torch.no_grad():
            x = b.next_observations
            B = x.shape[0]
            flat_x = x.view(B, 1, -1).expand(B, self.n, x.shape[1]).reshape(-1, x.shape[1])
            latent = noise.normal(B * self.n, 2 * self.act).clamp(-0.5, 0.5)
            sampled = vae_decode(self.imitator, flat_x, latent)
            actions = residual_policy(self.targ_pi, flat_x, sampled, self.flex)
            vals = q_continuous(self.targ_q, flat_x, actions, "none")  # (E, B*N, 1)
            E = vals.shape[0]
            values = vals.view(E, B, self.n, 1).transpose(0, 1)  # (B,E,N,1)
            mean_values = values.mean(dim=3)
            max_values, max_idx = mean_values.max(dim=1)
            min_values, min_idx = mean_values.min(dim=1)
            mix = (1.0 - self.lam) * max_values + self.lam * min_values
            action_idx = mix.argmax(dim=1)
            flat_values = values.transpose(1, 2).reshape(B * self.n, E, -1)
            bn = torch.arange(B * self.n)
            mx = flat_values[bn, max_idx.reshape(-1)].view(B, self.n, -1)
            mn = flat_values[bn, min_idx.reshape(-1)].view(B, self.n, -1)
            mix_values = (1.0 - self.lam) * mx + self.lam * mn
            return mix_values[torch.arange(B), action_idx]

    def update_critic(self, b, noise):
        self.critic_optim.zero_grad()
        q_tpn = self.compute_target(b, noise)
        loss = td_error_continuous(self.q, b.observations, b.actions, b.rewards, q_tpn, b.terminals,
                                   self.gamma ** b.n_steps)
        loss.backward()
        self.critic_optim.step()
        return float(loss.detach())

    def compute_actor_loss(self, b, noise):
        latent = noise.normal(b.observations.shape[0], 2 * self.act).clamp(-0.5, 0.5)
        sampled = vae_decode(self.imitator, b.observations, latent)
        action = residual_policy(self.pi, b.observations, sampled, self.flex)
        return -q_continuous(self.q, b.observations, action, "none")[0].mean()

    def update_actor(self, b, noise):
        self.actor_optim.zero_grad()
        loss = self.compute_actor_loss(b, noise)
        loss.backward()
        for v in list(self.q.values()) + list(self.imitator.values()):
            v.grad = None
        self.actor_optim.step()
        return float(loss.detach())

    def _update(self, b, noise):
        m = {"imitator_loss": self.update_imitator(b, noise)}
        if self.grad_step >= self.rl_start_step:
            m["critic_loss"] = self.update_critic(b, noise)
            if self.grad_step % self.update_actor_interval == 0:
                m["actor_loss"] = self.update_actor(b, noise)
                soft_sync(self.targ_pi, self.pi, self.tau)
                soft_sync(self.targ_q, self.q, self.tau)
        return m


class DiscreteCQL(_Algo):
    """DQN._update (algos/dqn.py:127-132) over DiscreteCQLImpl/DoubleDQNImpl
    (algos/torch/cql_impl.py:279-302, dqn_impl.py:97-171)."""

    def __init__(self, obs_shape, act, n_critics=1, lr=6.25e-5, gamma=0.99, target_update_interval=8000,
                 alpha=1.0, seed=0, critics=None, hidden=None, n_quantiles=None, double=True, conservative=True,
                 feature_size=512):
        """n_quantiles: QRQFunctionFactory members; double=False: DQNImpl.compute_target (dqn_impl.py:133-141);
        conservative=False: DQNImpl.compute_loss (dqn_impl.py:113-131)."""
        gen = torch.Generator().manual_seed(seed)
        self.nq, self.double, self.conservative = n_quantiles, double, conservative
        self.q = clone_params(critics if critics is not None
                              else make_discrete_critics(tuple(obs_shape), act, n_critics, gen, hidden, feature_size,
                                                         n_quantiles))
        self.targ_q = clone_params(self.q, False)
        self.optim = make_adam(self.q, lr)
        self.gamma, self.interval, self.alpha, self.act = gamma, target_update_interval, alpha, act
        self.grad_step = 0

    def compute_target(self, b):
        with torch.no_grad():
            sel = self.q if self.double else self.targ_q
            action = q_discrete(sel, b.next_observations, n_quantiles=self.nq).argmax(dim=1)
            if self.nq:
                th = quantiles_discrete(self.targ_q, b.next_observations, self.nq)  # (E,B,A,n)
                one_hot = F.one_hot(action.view(-1), num_classes=self.act).view(1, -1, self.act, 1).float()
                return reduce_quantile_ensemble_min((th * one_hot).sum(dim=2))  # (B,n)
            vals = q_discrete(self.targ_q, b.next_observations, "none")  # (E,B,A)
            one_hot = F.one_hot(action.view(-1), num_classes=self.act).float()
            picked = (vals * one_hot.unsqueeze(0)).sum(dim=2, keepdim=True)  # pick_value_by_action per member
            return reduce_ensemble(picked, "min")

    def compute_loss(self, b, q_tpn):
        act_long = b.actions.long()
        if self.nq:
            loss = td_error_discrete_qr(self.q, b.observations, act_long, b.rewards, q_tpn, b.terminals,
                                        self.gamma ** b.n_steps, self.nq)
        else:
            loss = td_error_discrete(self.q, b.observations, act_long, b.rewards, q_tpn, b.terminals,
                                     self.gamma ** b.n_steps)
        if not self.conservative:
            return loss
        policy_values = q_discrete(self.q, b.observations, n_quantiles=self.nq)
        lse = torch.logsumexp(policy_values, dim=1, keepdim=True)
        one_hot = F.one_hot(act_long.view(-1), num_classes=self.act)
        data_values = (q_discrete(self.q, b.observations, n_quantiles=self.nq) * one_hot).sum(dim=1, keepdim=True)
        return loss + self.alpha * (lse - data_values).mean()

    def _update(self, b, noise=None):
        self.optim.zero_grad()
        loss = self.compute_loss(b, self.compute_target(b))
        loss.backward()
        self.optim.step()
        if self.grad_step % self.interval == 0:
            hard_sync(self.targ_q, self.q)
        return {"loss": float(loss.detach())}


class DiscreteBCQ(DiscreteCQL):
    """DiscreteBCQ (algos/bcq.py:390-420 `_update` = DQN's) over DiscreteBCQImpl (algos/torch/bcq_impl.py:228-330): Double
    DQN whose greedy action is restricted to actions the behaviour-cloning head finds plausible
    (log pi(a|s) - max_a log pi > log action_flexibility), plus that head's loss (NLL + beta * mean(logits^2)) in the same
    objective and the same Adam.  Vector observations: the imitator has its own encoder (bcq_impl.py:276-282).
    Oracle only: the CUDA path is not built yet."""

    def __init__(self, obs_shape, act, n_critics=1, lr=6.25e-5, gamma=0.99, target_update_interval=8000,
                 action_flexibility=0.3, beta=0.5, seed=0, critics=None, imitator=None, hidden=None):
        super().__init__(obs_shape, act, n_critics=n_critics, lr=lr, gamma=gamma,
                         target_update_interval=target_update_interval, seed=seed, critics=critics, hidden=hidden,
                         double=True, conservative=False)
        gen = torch.Generator().manual_seed(seed + 1)
        if imitator is None:
            h = hidden or [256, 256]
            imitator = make_mlp("_encoder.", obs_shape[0], h, gen)
            imitator.update(make_head("_fc", act, h[-1], gen))
        self.imitator = clone_params(imitator)
        self.flex, self.beta = action_flexibility, beta
        self.optim = make_adam(OrderedDict(list(self.q.items()) + [("im." + k, v) for k, v in self.imitator.items()]), lr)

    def _logits(self, x):
        return F.linear(mlp_forward(self.imitator, "_encoder.", x), self.imitator["_fc.weight"], self.imitator["_fc.bias"])

    def best_action(self, x):
        log_probs = F.log_softmax(self._logits(x), dim=1)
        ratio = log_probs - log_probs.max(dim=1, keepdim=True).values
        mask = (ratio > math.log(self.flex)).float()
        value = q_discrete(self.q, x)
        normalized = value - value.min(dim=1, keepdim=True).values
        return (normalized * mask).argmax(dim=1)

    def compute_target(self, b):
        with torch.no_grad():
            action = self.best_action(b.next_observations)
            vals = q_discrete(self.targ_q, b.next_observations, "none")
            one_hot = F.one_hot(action.view(-1), num_classes=self.act).float()
            return reduce_ensemble((vals * one_hot.unsqueeze(0)).sum(dim=2, keepdim=True), "min")

    def compute_loss(self, b, q_tpn):
        loss = super().compute_loss(b, q_tpn)
        logits = self._logits(b.observations)
        imitator_loss = F.nll_loss(F.log_softmax(logits, dim=1), b.actions.long().view(-1)) + self.beta * (logits ** 2).mean()
        return loss + imitator_loss


class DiscreteSAC(_Algo):
    """DiscreteSAC._update (algos/sac.py:373-392) over DiscreteSACImpl (algos/torch/sac_impl.py:165-420): categorical
    policy, expectation-form soft target sum_a pi(a|s') (min_e Q'_e(s', a) - temp * log pi(a|s')), Huber critics,
    temperature towards 0.98 * log |A|, hard target copy every `target_update_interval` steps; every Adam with
    eps = 1e-4 (algos/sac.py:305-307).  `log_probs` is `Categorical(softmax(h)).logits`, i.e. log of the re-normalised,
    eps-clamped probabilities (policies.py:303-306,350-352).  Oracle only: the CUDA path is not built yet."""

    def __init__(self, obs, act, hidden=(256, 256), n_critics=2, actor_lr=3e-4, critic_lr=3e-4, temp_lr=3e-4, gamma=0.99,
                 initial_temperature=1.0, target_update_interval=8000, adam_eps=1e-4, seed=0, policy=None, critics=None):
        gen = torch.Generator().manual_seed(seed)
        self.q = clone_params(critics if critics is not None
                              else make_discrete_critics((obs,), act, n_critics, gen, list(hidden)))
        if policy is None:
            policy = make_mlp("_encoder.", obs, hidden, gen)
            policy.update(make_head("_fc", act, hidden[-1], gen))
        self.pi = clone_params(policy)
        self.targ_q = clone_params(self.q, False)
        self.log_temp = {"_parameter": torch.full((1, 1), math.log(initial_temperature)).requires_grad_(True)}
        self.critic_optim = make_adam(self.q, critic_lr, eps=adam_eps)
        self.actor_optim = make_adam(self.pi, actor_lr, eps=adam_eps)
        self.temp_optim = make_adam(self.log_temp, temp_lr, eps=adam_eps)
        self.temp_lr, self.gamma, self.interval, self.act = temp_lr, gamma, target_update_interval, act
        self.grad_step = 0

    def log_probs(self, x):
        h = F.linear(mlp_forward(self.pi, "_encoder.", x), self.pi["_fc.weight"], self.pi["_fc.bias"])
        return torch.distributions.Categorical(torch.softmax(h, dim=1)).logits

    def update_temp(self, b):
        self.temp_optim.zero_grad()
        with torch.no_grad():
            log_probs = self.log_probs(b.observations)
            expct = (log_probs.exp() * log_probs).sum(dim=1, keepdim=True)
            targ_temp = expct + 0.98 * (-math.log(1 / self.act))
        loss = -(self.log_temp["_parameter"].exp() * targ_temp).mean()
        loss.backward()
        self.temp_optim.step()
        return float(loss.detach()), float(self.log_temp["_parameter"].exp().detach()[0][0])

    def compute_target(self, b):
        with torch.no_grad():
            log_probs = self.log_probs(b.next_observations)
            entropy = self.log_temp["_parameter"].exp() * log_probs
            target = q_discrete(self.targ_q, b.next_observations, "min")
            return (log_probs.exp() * (target - entropy)).sum(dim=1, keepdim=True)

    def compute_actor_loss(self, b):
        with torch.no_grad():
            q_t = q_discrete(self.q, b.observations, "min")
        log_probs = self.log_probs(b.observations)
        entropy = self.log_temp["_parameter"].exp() * log_probs
        return (log_probs.exp() * (entropy - q_t)).sum(dim=1).mean()

    def _update(self, b, noise=None):
        m = {}
        if self.temp_lr > 0:
            m["temp_loss"], m["temp"] = self.update_temp(b)
        self.critic_optim.zero_grad()
        c_loss = td_error_discrete(self.q, b.observations, b.actions.long(), b.rewards, self.compute_target(b),
                                   b.terminals, self.gamma ** b.n_steps)
        c_loss.backward()
        self.critic_optim.step()
        self.actor_optim.zero_grad()
        a_loss = self.compute_actor_loss(b)
        a_loss.backward()
        self.log_temp["_parameter"].grad = None
        self.actor_optim.step()
        if self.grad_step % self.interval == 0:
            hard_sync(self.targ_q, self.q)
        m["critic_loss"], m["actor_loss"] = float(c_loss.detach()), float(a_loss.detach())
        return m


class TD3PlusRelation(TD3PlusBC):
    """TD3PlusRelation._update (algos/td3_relational.py:176-192) over TD3PlusRelationImpl
    (algos/torch/td3_relational_impl.py:17-96) -- the algorithm this fork adds: TD3+BC's schedule and lambda-normalised
    -Q term, with the cloning term replaced by a batch-relational distillation loss between the B x B similarity
    matrices  a_data a_data^T / 0.04 (softmax, detached)  and  pi(s) a_data^T / 0.1 (log-softmax).  Extra metrics
    absQmean / RelationLoss / TD3Loss / BCLoss are those of the most recent actor step.  Oracle only."""

    T_K, T_Q = 0.04, 0.1

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.log_metrics: Dict[str, float] = {}

    def compute_actor_loss(self, b: Batch):
        action = deterministic_policy(self.pi, b.observations)
        q_t = q_continuous(self.q, b.observations, action, "none")[0]
        lam = self.alpha / (q_t.abs().mean()).detach()
        logits_q = torch.einsum("nc,kc->nk", [action, b.actions])
        logits_k = torch.einsum("nc,kc->nk", [b.actions, b.actions])
        relation = -torch.sum(F.softmax(logits_k.detach() / self.T_K, dim=1) * F.log_softmax(logits_q / self.T_Q, dim=1),
                              dim=1).mean()
        self.log_metrics = {"absQmean": float(q_t.abs().mean().detach()), "RelationLoss": float(relation.detach()),
                            "TD3Loss": float(-q_t.mean().detach()),
                            "BCLoss": float(((b.actions - action.detach()) ** 2).mean())}
        return lam * -q_t.mean() + relation + False * ((b.actions - action) ** 2).mean()

    def _update(self, b, noise):
        m = super()._update(b, noise)
        m.update(self.log_metrics)
        return m


class BC(_Algo):
    """BC / DiscreteBC `_update` (algos/bc.py:58-61) over BCImpl / DiscreteBCImpl (algos/torch/bc_impl.py:24-242):
    one Adam step (lr 1e-3) on DeterministicRegressor.compute_error = mse(tanh(fc(enc(s))), a) (imitators.py:158-175),
    or for discrete actions on DiscreteImitator.compute_error = nll(log_softmax(logits), a) + beta * mean(logits^2)
    (imitators.py:136-160).  Oracle only: the CUDA path is not built yet."""

    def __init__(self, obs, act, hidden=(256, 256), lr=1e-3, discrete=False, beta=0.5, seed=0, imitator=None):
        gen = torch.Generator().manual_seed(seed)
        if imitator is None:
            imitator = make_mlp("_encoder.", obs, hidden, gen)
            imitator.update(make_head("_fc", act, hidden[-1], gen))
        self.imitator = clone_params(imitator)
        self.optim = make_adam(self.imitator, lr)
        self.discrete, self.beta = discrete, beta
        self.grad_step = 0

    def _out(self, x):
        return F.linear(mlp_forward(self.imitator, "_encoder.", x), self.imitator["_fc.weight"], self.imitator["_fc.bias"])

    def compute_loss(self, b: Batch):
        if self.discrete:
            logits = self._out(b.observations)
            return F.nll_loss(F.log_softmax(logits, dim=1), b.actions.long().view(-1)) + self.beta * (logits ** 2).mean()
        return F.mse_loss(torch.tanh(self._out(b.observations)), b.actions)

    def predict(self, x):
        with torch.no_grad():
            out = self._out(x)
            return out.argmax(dim=1) if self.discrete else torch.tanh(out)

    def _update(self, b, noise=None):
        self.optim.zero_grad()
        loss = self.compute_loss(b)
        loss.backward()
        self.optim.step()
        return {"loss": float(loss.detach())}
