"""GPU: each CUDA kernel, called through the C ABI, against a plain fp32/fp64 PyTorch statement of
the same op.  Tolerance for fp32 mode: 1e-5 relative (north_star)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _dev():
    return torch.device("cuda:0")


def _st():
    return torch.cuda.current_stream().cuda_stream


def _close(a, b, rtol=RTOL, atol=None, msg=""):
    b = b.to(a.dtype)
    scale = max(1.0, float(b.abs().max()))
    err = float((a - b).abs().max())
    assert err <= (atol if atol is not None else rtol * scale), f"{msg} max err {err} scale {scale}"


@pytest.mark.parametrize("M,N,K,E,shared", [(256, 256, 23, 2, True), (7936, 256, 256, 2, False), (33, 70, 19, 3, False),
                                            (512, 300, 400, 1, False), (1000, 256, 256, 10, False)])
def test_linear_forward_backward(M, N, K, E, shared):
    from d3rlpy_b200._lib import lib

    L, dev = lib(), _dev()
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    x = torch.randn(1 if shared else E, M, K, generator=g).to(dev)
    w = (torch.randn(E, N, K, generator=g) / math.sqrt(K)).to(dev)
    b = torch.randn(E, N, generator=g).to(dev)
    y = torch.empty(E, M, N, device=dev)
    sx = 0 if shared else M * K
    L.linear_forward(x.data_ptr(), K, sx, w.data_ptr(), K, N * K, b.data_ptr(), N, y.data_ptr(), N, M * N, M, N, K, E,
                     1, _st())
    xe = x.expand(E, M, K).double()
    ref = torch.relu(torch.einsum("emk,enk->emn", xe, w.double()) + b.double()[:, None, :])
    _close(y, ref, msg="forward")

    dy = torch.randn(E, M, N, generator=g).to(dev)
    dx = torch.empty(E, M, K, device=dev)
    src = torch.randn(E, M, K, generator=g).to(dev)
    L.linear_backward_data(dy.data_ptr(), N, M * N, w.data_ptr(), K, N * K, dx.data_ptr(), K, M * K, src.data_ptr(),
                           K, M * K, M, N, K, E, _st())
    ref = torch.einsum("emn,enk->emk", dy.double(), w.double()) * (src > 0).double()
    _close(dx, ref, msg="dgrad")

    # column-restricted dgrad without mask (action columns of layer 1)
    if K > 8:
        c0, nc = K - 6, 6
        dxa = torch.empty(E, M, nc, device=dev)
        L.linear_backward_data(dy.data_ptr(), N, M * N, w.data_ptr() + 4 * c0, K, N * K, dxa.data_ptr(), nc, M * nc,
                               None, 0, 0, M, N, nc, E, _st())
        ref = torch.einsum("emn,enk->emk", dy.double(), w.double()[:, :, c0:c0 + nc])
        _close(dxa, ref, msg="dgrad cols")

    dw = torch.zeros(E, N, K, device=dev)
    db = torch.zeros(E, N, device=dev)
    L.linear_backward_weight(dy.data_ptr(), N, M * N, x.data_ptr(), K, sx, dw.data_ptr(), K, N * K, db.data_ptr(), N,
                             M, N, K, E, _st())
    refw = torch.einsum("emn,emk->enk", dy.double(), xe)
    _close(dw, refw, rtol=2e-5, msg="wgrad")
    _close(db, dy.double().sum(1), rtol=2e-5, msg="bgrad")


@pytest.mark.parametrize("M,N,K,E,tanh", [(7936, 1, 256, 2, False), (512, 12, 256, 1, False), (100, 6, 300, 1, True),
                                          (77, 24, 750, 1, False)])
def test_head_forward_backward(M, N, K, E, tanh):
    from d3rlpy_b200._lib import lib

    L, dev = lib(), _dev()
    g = torch.Generator().manual_seed(N * 7 + K)
    x = torch.relu(torch.randn(E, M, K, generator=g)).to(dev)
    w = (torch.randn(E, N, K, generator=g) / math.sqrt(K)).to(dev)
    b = torch.randn(E, N, generator=g).to(dev)
    y = torch.empty(E, M, N, device=dev)
    L.head_forward(x.data_ptr(), K, M * K, w.data_ptr(), K, N * K, b.data_ptr(), N, y.data_ptr(), N, M * N, M, N, K, E,
                   int(tanh), _st())
    ref = torch.einsum("emk,enk->emn", x.double(), w.double()) + b.double()[:, None, :]
    if tanh:
        ref = torch.tanh(ref)
    _close(y, ref, msg="head fwd")
    dy = torch.randn(E, M, N, generator=g).to(dev)
    dx = torch.empty(E, M, K, device=dev)
    L.head_backward_data(dy.data_ptr(), N, M * N, w.data_ptr(), K, N * K, dx.data_ptr(), K, M * K, x.data_ptr(), K,
                         M * K, M, N, K, E, _st())
    ref = torch.einsum("emn,enk->emk", dy.double(), w.double()) * (x > 0).double()
    _close(dx, ref, msg="head dgrad")
    dw = torch.zeros(E, N, K, device=dev)
    db = torch.zeros(E, N, device=dev)
    L.head_backward_weight(dy.data_ptr(), N, M * N, x.data_ptr(), K, M * K, dw.data_ptr(), K, N * K, db.data_ptr(), N,
                           M, N, K, E, _st())
    _close(dw, torch.einsum("emn,emk->enk", dy.double(), x.double()), rtol=2e-5, msg="head wgrad")
    _close(db, dy.double().sum(1), rtol=2e-5, msg="head bgrad")


@pytest.mark.parametrize("n", [1, 7, 4096, 415_248])
def test_adam_and_soft_sync_match_torch(n):
    """params, exp_avg, exp_avg_sq after 3 steps vs torch.optim.Adam (CPU, the reference's optimizer),
    and target vs soft_sync's closed form (tests/test_torch_utility.py:30-57 of the reference)."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), _dev()
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=3e-4)
    targ_ref = torch.randn(n, generator=g)
    p = p0.clone().to(dev)
    targ = targ_ref.clone().to(dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    tau = 0.005
    for it in range(3):
        grad = torch.randn(n, generator=g) * (10.0 ** (it - 1))
        p_ref.grad = grad.clone()
        opt.step()
        with torch.no_grad():
            targ_ref.mul_(1 - tau)
            targ_ref.add_(tau * p_ref.data)
        gd = grad.to(dev)
        L.tick(step.data_ptr(), 1, 1, _st())
        L.adam_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), targ.data_ptr(), n, step.data_ptr(), 3e-4,
                    0.9, 0.999, 1e-8, tau, 1, _st())
        assert float(gd.abs().max()) == 0.0  # grads are cleared for the next accumulation
    st = opt.state[p_ref]
    torch.testing.assert_close(p.cpu(), p_ref.data, rtol=1e-6, atol=1e-7)
    # moments: 1e-5 relative (fp32-mode tolerance); the CPU lerp/addcmul differ from ours only by FMA contraction
    torch.testing.assert_close(m.cpu(), st["exp_avg"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(v.cpu(), st["exp_avg_sq"], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(targ.cpu(), targ_ref, rtol=1e-6, atol=1e-7)
    # stand-alone soft_sync and hard_sync
    t2 = torch.randn(n, generator=g)
    t2d = t2.clone().to(dev)
    L.soft_sync(t2d.data_ptr(), p.data_ptr(), n, tau, _st())
    expect = t2 * (1 - tau) + tau * p.cpu()
    torch.testing.assert_close(t2d.cpu(), expect, rtol=1e-6, atol=1e-7)
    L.hard_sync(t2d.data_ptr(), p.data_ptr(), n, _st())
    assert torch.equal(t2d, p)


def test_policy_sample_rows_matches_squashed_gaussian():
    """tanh-Gaussian sample + log-prob vs the oracle's restatement of SquashedGaussianDistribution
    (reference test: tests/models/torch/test_distributions.py:51-96, atol 1e-2; here 1e-5)."""
    from d3rlpy_b200._lib import lib
    from oracle import update as ou

    L, dev = lib(), _dev()
    B, N, O, A = 37, 5, 9, 6
    g = torch.Generator().manual_seed(5)
    head = torch.randn(B, 2 * A, generator=g) * 2.0
    head[0, A:] = 5.0    # clamps at max_logstd
    head[1, A:] = -30.0  # clamps at min_logstd
    eps = torch.randn(N, B, A, generator=g)
    obs = torch.randn(B, O, generator=g)
    x = torch.zeros(B * N, O + A, device=dev)
    lp = torch.zeros(B * N, device=dev)
    head_d, eps_d, obs_d = head.to(dev), eps.to(dev), obs.to(dev)  # keep the device copies alive
    L.policy_sample_rows(head_d.data_ptr(), 2 * A, eps_d.data_ptr(), obs_d.data_ptr(), O,
                         x.data_ptr(), O + A, None, lp.data_ptr(), B, N, O, A, -20.0, 2.0, 0, _st())
    torch.cuda.synchronize()
    mu, std = head[:, :A], head[:, A:].clamp(-20, 2).exp()
    raw = mu.unsqueeze(0) + eps * std.unsqueeze(0)
    ref_a = torch.tanh(raw).transpose(0, 1).reshape(B * N, A)
    ref_lp = ou.squashed_log_prob(mu.unsqueeze(0), std.unsqueeze(0), raw).transpose(0, 1).reshape(B * N)
    _close(x[:, O:].cpu(), ref_a, msg="action")
    _close(x[:, :O].cpu(), obs.repeat_interleave(N, 0), atol=0.0, msg="obs")
    _close(lp.cpu(), ref_lp, rtol=2e-5, msg="logp")


def test_critic_loss_and_gradient_vs_autograd():
    """TD (sum over members of batch means) + CQL conservative term and dL/dQ vs torch autograd."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), _dev()
    B, N, A, E = 19, 4, 3, 3
    R = B * (1 + 3 * N)
    g = torch.Generator().manual_seed(11)
    q = (torch.randn(E, R, generator=g) * 3).requires_grad_(True)
    qt = torch.randn(E, B, generator=g)
    rew, term = torch.randn(B, generator=g), (torch.rand(B, generator=g) < 0.3).float()
    ns = torch.randint(1, 4, (B,), generator=g).float()
    lp1, lp2 = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    log_alpha = torch.tensor([0.3])
    gamma, cw, thr = 0.99, 5.0, 10.0
    y = rew + (gamma ** ns) * qt.min(0).values * (1 - term)
    td = sum(F.mse_loss(q[e, :B], y, reduction="none").mean() for e in range(E))
    v1 = q[:, B:B + B * N].view(E, B, N) - lp1
    v2 = q[:, B + B * N:B + 2 * B * N].view(E, B, N) - lp2
    v3 = q[:, B + 2 * B * N:].view(E, B, N) - math.log(0.5 ** A)
    lse = torch.logsumexp(torch.cat([v1, v2, v3], 2), 2, keepdim=True)
    cons = log_alpha.exp().clamp(0, 1e6)[0] * (cw * (lse.mean(0).mean() - q[:, :B].mean(0).mean()) - thr)
    loss = td + cons
    loss.backward()
    qd = q.detach().to(dev)
    dq = torch.zeros(E, R, device=dev)
    sums = torch.zeros(4, device=dev)
    metric = torch.zeros(2, device=dev)
    la = log_alpha.to(dev)
    lps = torch.stack([lp1.reshape(-1), lp2.reshape(-1)]).to(dev)
    d = lambda t: t.to(dev).data_ptr()
    keep = [qt.to(dev), rew.to(dev), term.to(dev), ns.to(dev)]
    L.critic_loss(qd.data_ptr(), R, keep[0].data_ptr(), B, E, None, keep[1].data_ptr(), keep[2].data_ptr(),
                  keep[3].data_ptr(), gamma, lps.data_ptr(), lps.data_ptr() + 4 * B * N, N, A, la.data_ptr(), cw,
                  dq.data_ptr(), R, sums.data_ptr(), None, B, E, 1.0 / B, 1, _st())
    L.cql_finalize(sums.data_ptr(), la.data_ptr(), 1.0 / B, E, cw, thr, 0, 1, metric.data_ptr(), None, _st())
    _close(metric[0].cpu(), loss.detach(), msg="loss")
    _close(dq.cpu(), q.grad, rtol=2e-5, msg="dq")
    # alpha mode: loss = -cons, d/dlog_alpha
    la_t = log_alpha.clone().requires_grad_(True)
    cons2 = la_t.exp().clamp(0, 1e6)[0] * (cw * (lse.detach().mean(0).mean() - q.detach()[:, :B].mean(0).mean()) - thr)
    (-cons2).backward()
    L.cql_finalize(sums.data_ptr(), la.data_ptr(), 1.0 / B, E, cw, thr, 1, 1, metric.data_ptr(),
                   metric.data_ptr() + 4, _st())
    _close(metric[0].cpu(), -cons2.detach(), msg="alpha loss")
    _close(metric[1].cpu(), la_t.grad[0], msg="alpha grad")


def test_noise_fill_statistics():
    from d3rlpy_b200._lib import lib

    L, dev = lib(), _dev()
    nn_, nu = 1_000_001, 500_003
    out = torch.zeros(nn_ + nu, device=dev)
    ctr = torch.zeros(1, dtype=torch.int32, device=dev)
    L.noise_fill(out.data_ptr(), nn_, nu, 1234, ctr.data_ptr(), _st())
    a = out[:nn_].double()
    u = out[nn_:].double()
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1) < 5e-3
    assert abs(float((a ** 4).mean()) - 3.0) < 0.1
    assert float(u.min()) >= -1 and float(u.max()) <= 1 and abs(float(u.mean())) < 5e-3
    assert abs(float(u.var()) - 1 / 3) < 5e-3
    first = out.clone()
    ctr += 1
    L.noise_fill(out.data_ptr(), nn_, nu, 1234, ctr.data_ptr(), _st())
    assert float((first == out).float().mean()) < 1e-3  # a new epoch gives a new stream


# ----------------------------------------------------------------------------------------- fused CQL glue kernels
def test_cql_fused_glue_kernels_match_unfused_kernels():
    """csrc/cql_fused.cu vs the separately-verified kernels it replaces, on identical inputs:
    cql_rows == concat_rows + policy_sample_rows (+ bf16 rounding of the rows, log-probs to 1e-6);
    cql_loss_step == critic_loss + cql_finalize (+ scalar_adam in alpha mode); sac_temp_step == sac_temp_loss +
    scalar_adam; sac_actor_step == sac_actor_loss."""
    import ctypes

    from d3rlpy_b200._lib import lib

    L, dev, st = lib(), torch.device("cuda:0"), torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(0)
    B, N, O, A, E = 48, 5, 7, 3, 2
    R = B * (1 + 3 * N)
    ld = (O + A + 7) // 8 * 8
    rnd = lambda *s: torch.randn(*s, generator=g).to(dev)
    head = rnd(2 * B, 2 * A)
    obs, nobs = rnd(B, O), rnd(B, O)
    act = (torch.rand(B, A, generator=g) * 2 - 1).to(dev)
    eps = {k: rnd(N, B, A) for k in ("ct", "ctp1", "at", "atp1")}
    rand = {k: (torch.rand(B * N, A, generator=g) * 2 - 1).to(dev) for k in ("c", "a")}
    eps_soft, eps_actor, eps_temp = rnd(B, A), rnd(B, A), rnd(B, A)
    X = torch.zeros(2 * R + 2 * B, ld, dtype=torch.bfloat16, device=dev)
    lp = torch.zeros(4, B * N, device=dev)
    lpm = torch.zeros(3, B, device=dev)
    p = lambda t: t.data_ptr()
    ptrs = [p(eps["ct"]), p(eps["ctp1"]), p(rand["c"]), p(lp[0]), p(lp[1]), p(eps["at"]), p(eps["atp1"]), p(rand["a"]),
            p(lp[2]), p(lp[3]), p(eps_soft), p(lpm[0]), p(eps_actor), p(lpm[1]), p(eps_temp), p(lpm[2])]
    L.cql_rows(p(head), p(obs), p(nobs), p(act), B, N, O, A, -20.0, 2.0, p(X), ld, 2, (ctypes.c_void_p * 16)(*ptrs),
               (ctypes.c_int64 * 4)(0, R, 2 * R, 2 * R + B), st)
    # reference rows with the unfused kernels (fp32)
    def ref_group(e_t, e_tp1, rnd_a):
        x = torch.zeros(R, O + A, device=dev)
        l = torch.zeros(2, B * N, device=dev)
        L.concat_rows(p(obs), O, p(act), A, None, 0.0, 0.0, 0.0, p(x), O + A, B, 1, O, A, st)
        L.policy_sample_rows(p(head), 2 * A, p(e_t), p(obs), O, p(x) + 4 * (O + A) * B, O + A, None, p(l[0]), B, N, O, A,
                             -20.0, 2.0, 0, st)
        L.policy_sample_rows(p(head) + 4 * B * 2 * A, 2 * A, p(e_tp1), p(obs), O, p(x) + 4 * (O + A) * (B + B * N), O + A,
                             None, p(l[1]), B, N, O, A, -20.0, 2.0, 0, st)
        L.concat_rows(p(obs), O, p(rnd_a), A, None, 0.0, 0.0, 0.0, p(x) + 4 * (O + A) * (B + 2 * B * N), O + A, B, N, O, A, st)
        return x, l
    xc, lc = ref_group(eps["ct"], eps["ctp1"], rand["c"])
    xa, la_ = ref_group(eps["at"], eps["atp1"], rand["a"])
    xt = torch.zeros(B, O + A, device=dev)
    lsoft = torch.zeros(B, device=dev)
    L.policy_sample_rows(p(head) + 4 * B * 2 * A, 2 * A, p(eps_soft), p(nobs), O, p(xt), O + A, None, p(lsoft), B, 1, O, A,
                         -20.0, 2.0, 0, st)
    xact = torch.zeros(B, O + A, device=dev)
    lact, ltemp = torch.zeros(B, device=dev), torch.zeros(B, device=dev)
    L.policy_sample_rows(p(head), 2 * A, p(eps_actor), p(obs), O, p(xact), O + A, None, p(lact), B, 1, O, A, -20.0, 2.0, 0, st)
    L.policy_sample_rows(p(head), 2 * A, p(eps_temp), None, 0, None, 0, None, p(ltemp), B, 1, 0, A, -20.0, 2.0, 0, st)
    torch.cuda.synchronize()
    W = O + A
    assert torch.equal(X[:R, :W], xc.to(torch.bfloat16)) and torch.equal(X[R:2 * R, :W], xa.to(torch.bfloat16))
    assert torch.equal(X[2 * R:2 * R + B, :W], xt.to(torch.bfloat16))
    assert torch.equal(X[2 * R + B:, :W], xact.to(torch.bfloat16))
    # the fp32-mode variant writes the same rows as fp32 operands (padded leading dimension)
    ld4 = (W + 3) // 4 * 4
    X32 = torch.zeros(2 * R + 2 * B, ld4, device=dev)
    L.cql_rows_f32(p(head), p(obs), p(nobs), p(act), B, N, O, A, -20.0, 2.0, p(X32), ld4, 2, (ctypes.c_void_p * 16)(*ptrs),
                   (ctypes.c_int64 * 4)(0, R, 2 * R, 2 * R + B), st)
    torch.cuda.synchronize()
    for got, want, what in ((X32[:R], xc, "critic rows"), (X32[R:2 * R], xa, "alpha rows"),
                            (X32[2 * R:2 * R + B], xt, "target rows"), (X32[2 * R + B:], xact, "actor rows")):
        _close(got[:, :W], want, rtol=1e-6, msg=what)
    _close(lp[:2], lc, rtol=1e-6, msg="critic logp")
    _close(lp[2:], la_, rtol=1e-6, msg="alpha logp")
    _close(lpm[0], lsoft, rtol=1e-6, msg="soft logp")
    _close(lpm[1], lact, rtol=1e-6, msg="actor logp")
    _close(lpm[2], ltemp, rtol=1e-6, msg="temp logp")

    # ---- loss kernels
    q = rnd(E, R) * 3
    q_t = rnd(E, B)
    rew, term = rnd(B), (torch.rand(B, generator=g) < 0.2).float().to(dev)
    nst = torch.randint(1, 4, (B,), generator=g).float().to(dev)
    for mode in (0, 1):
        sc_a = torch.zeros(16, device=dev); sc_a[0] = 0.3
        sc_b = sc_a.clone()
        steps = torch.tensor([3], dtype=torch.int32, device=dev)
        sums_a, sums_b = torch.zeros(4, device=dev), torch.zeros(4, device=dev)
        dq_a, dq_b = torch.zeros(E, R, device=dev), torch.zeros(E, R, device=dev)
        met_a, met_b = torch.zeros(2, device=dev), torch.zeros(2, device=dev)
        done = torch.zeros(4 + 3 * ((B * E + 7) // 8), dtype=torch.int32, device=dev)   # counter + per-block partials
        td = mode == 0
        L.cql_loss_step(p(q), R, p(q_t) if td else None, B, E, None, p(rew) if td else None, p(term) if td else None,
                        p(nst) if td else None, 0.99, p(lc[0]), p(lc[1]), N, A, p(sc_a), 5.0, 10.0,
                        p(dq_a) if td else None, R, p(sums_a), p(done), B, E, 1.0 / B, mode, p(steps), 1e-4, p(met_a),
                        p(met_a) + 4, st)
        L.critic_loss(p(q), R, p(q_t) if td else None, B, E, None, p(rew), p(term), p(nst), 0.99, p(lc[0]), p(lc[1]), N, A,
                      p(sc_b), 5.0, p(dq_b) if td else None, R, p(sums_b), None, B, E, 1.0 / B, 1 if td else 0, st)
        L.cql_finalize(p(sums_b), p(sc_b), 1.0 / B, E, 5.0, 10.0, mode, 1, p(met_b), p(sc_b) + 16 if mode else None, st)
        if mode:
            L.scalar_adam(p(sc_b), p(sc_b) + 16, p(sc_b) + 32, p(sc_b) + 48, p(steps), 1e-4, 0.9, 0.999, 1e-8, p(met_b) + 4, st)
        torch.cuda.synchronize()
        _close(met_a[:1 + mode], met_b[:1 + mode], rtol=2e-6, msg=f"loss metric mode {mode}")
        _close(sc_a, sc_b, rtol=1e-6, msg="log_alpha adam state")
        if td:
            _close(dq_a, dq_b, rtol=2e-6, msg="dq")
        assert int(done[0]) == 0
    # ---- temp / actor
    sc_a = torch.zeros(16, device=dev); sc_a[0] = -0.2
    sc_b = sc_a.clone()
    steps = torch.tensor([2], dtype=torch.int32, device=dev)
    met_a, met_b = torch.zeros(2, device=dev), torch.zeros(2, device=dev)
    L.sac_temp_step(p(ltemp), p(sc_a), p(steps), B, A, 1.0 / B, 1e-4, p(met_a), p(met_a) + 4, st)
    L.sac_temp_loss(p(ltemp), p(sc_b), B, A, 1.0 / B, p(met_b), p(sc_b) + 16, 0, st)
    L.scalar_adam(p(sc_b), p(sc_b) + 16, p(sc_b) + 32, p(sc_b) + 48, p(steps), 1e-4, 0.9, 0.999, 1e-8, p(met_b) + 4, st)
    qa = rnd(E, B)
    dqa, dqb = torch.zeros(E, B, device=dev), torch.zeros(E, B, device=dev)
    ls_a, ls_b = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    done = torch.zeros(4 + (B + 255) // 256, dtype=torch.int32, device=dev)
    m_act = torch.zeros(1, device=dev)
    L.sac_actor_step(p(qa), B, p(lact), p(sc_a), p(dqa), B, p(ls_a), p(done), p(m_act), B, E, 1.0 / B, st)
    L.sac_actor_loss(p(qa), B, p(lact), p(sc_a), p(dqb), B, p(ls_b), B, E, 1.0 / B, st)
    torch.cuda.synchronize()
    _close(met_a, met_b, rtol=1e-6, msg="temp metrics")
    _close(sc_a, sc_b, rtol=1e-6, msg="log_temp adam state")
    assert torch.equal(dqa, dqb)
    _close(m_act, ls_b, rtol=1e-6, msg="actor loss")


# ----------------------------------------------------------------------------------------- quantile-regression head
@pytest.mark.parametrize("B,A,NQ,E,conservative", [(64, 6, 200, 3, 1), (33, 18, 32, 1, 1), (5, 2, 1, 2, 0),
                                                   (256, 4, 51, 2, 0)])
def test_qr_target_loss_values_vs_autograd(B, A, NQ, E, conservative):
    """csrc/qr.cu through the C ABI vs the oracle's statement of qr_q_function.py / utility.py:35-61 with autograd."""
    from d3rlpy_b200._lib import lib
    from oracle import update as ou

    L, dev = lib(), _dev()
    g = torch.Generator().manual_seed(B * 7 + NQ)
    th_on = torch.randn(E, B, A, NQ, generator=g)
    th_tg = torch.randn(E, B, A, NQ, generator=g)
    act = torch.randint(A, (B,), generator=g)
    rew, term = torch.randn(B, 1, generator=g), (torch.rand(B, 1, generator=g) < 0.2).float()
    nsteps = torch.randint(1, 4, (B, 1), generator=g).float()
    gamma, alpha = 0.99, 0.7

    # ---- target: greedy action of the member-mean values of `select`, quantiles of the min-mean target member
    a_star = th_on.mean(dim=3).mean(dim=0).argmax(dim=1)
    one_hot = F.one_hot(a_star, A).view(1, B, A, 1).float()
    ref_tpn = ou.reduce_quantile_ensemble_min((th_tg * one_hot).sum(dim=2))
    d_on, d_tg = th_on.to(dev), th_tg.to(dev)
    q_tpn = torch.empty(B, NQ, device=dev)
    L.qr_target(d_on.data_ptr(), B * A * NQ, d_tg.data_ptr(), B * A * NQ, q_tpn.data_ptr(), B, A, NQ, E, _st())
    assert torch.equal(q_tpn.cpu(), ref_tpn)

    # ---- values
    vals = torch.empty(E, B, A, device=dev)
    L.qr_values(d_on.data_ptr(), B * A * NQ, vals.data_ptr(), B * A, B, A, NQ, E, _st())
    _close(vals.cpu(), th_on.mean(dim=3), msg="values")

    # ---- loss + gradient
    th = th_on.clone().double().requires_grad_(True)
    taus = ou.make_taus(NQ).double()
    oh = F.one_hot(act, A).view(B, A, 1).double()
    td = torch.zeros((), dtype=torch.float64)
    for e in range(E):
        picked = (th[e] * oh).sum(dim=1)
        td = td + ou.quantile_huber_loss(picked, rew.double(), ref_tpn.double(), term.double(), taus,
                                         (gamma ** nsteps).double()).mean()
    cons = torch.zeros((), dtype=torch.float64)
    if conservative:
        v = th.mean(dim=3).mean(dim=0)
        cons = (torch.logsumexp(v, dim=1) - (v * F.one_hot(act, A)).sum(dim=1)).mean()
    total = td + alpha * cons
    total.backward()
    sums, partials = torch.zeros(2, device=dev), torch.empty(2 * B, device=dev)
    dth = torch.full((E, B, A, NQ), float("nan"), device=dev)
    metric = torch.zeros(1, device=dev)
    t = lambda x: x.to(dev).contiguous()
    d_act, d_rew, d_term, d_ns = t(act.float()), t(rew), t(term), t(nsteps)
    L.qr_loss(d_on.data_ptr(), B * A * NQ, q_tpn.data_ptr(), d_act.data_ptr(), d_rew.data_ptr(), d_term.data_ptr(),
              d_ns.data_ptr(), gamma, alpha, dth.data_ptr(), B * A * NQ, partials.data_ptr(), sums.data_ptr(), B, A, NQ, E, 1.0 / B,
              conservative, _st())
    L.dcql_finalize(sums.data_ptr(), 1.0 / B, alpha, conservative, metric.data_ptr(), _st())
    torch.cuda.synchronize()
    assert abs(float(metric) - float(total)) <= 2e-5 * max(1.0, abs(float(total))), (float(metric), float(total))
    ref_g = th.grad
    err = float((dth.cpu().double() - ref_g).abs().max())
    assert err <= 1e-5 * max(float(ref_g.abs().max()), 1e-12) + 1e-9, (err, float(ref_g.abs().max()))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,feat,n,E", [(32, 64, 96, 2), (100, 256, 800, 1), (7, 48, 40, 3)])
def test_wide_head_dense_layer_forward_backward(precision, rows, feat, n, E):
    """nets.DenseNet with a head wider than 32 outputs (quantile heads): forward and backward vs autograd."""
    from d3rlpy_b200.nets import DenseNet

    dev = _dev()
    O = 10
    g = torch.Generator().manual_seed(rows + n)
    net = DenseNet(O, [feat], [("_fc", n)], E, dev, trunk_prefix="_encoder.", member_key="_q_funcs.{e}.{name}",
                   seed_gen=g, precision=precision)
    assert net.wide_head and not net.fused_ok
    st = _st()
    net.refresh_shadow("params", st)
    x = torch.randn(rows, O, generator=g).to(dev)
    ctx = net.ctx("t", rows, E, True)
    out = torch.empty(E, rows, n, device=dev)
    net.forward("params", x, O, rows, ctx, out, st)
    d_out = (torch.randn(E, rows, n, generator=g) / n).to(dev)
    net.arena.grads.zero_()
    net.backward(x, O, rows, ctx, d_out, st)
    torch.cuda.synchronize()
    tol = 1e-5 if precision == "fp32" else 3e-2
    for e in range(E):
        p = {k: net.arena.view(k, e).detach().clone().cpu().double().requires_grad_(True)
             for k in ("_encoder._fcs.0.weight", "_encoder._fcs.0.bias", "__head.weight", "__head.bias")}
        pre = F.linear(x.cpu().double(), p["_encoder._fcs.0.weight"], p["_encoder._fcs.0.bias"])
        if precision == "bf16":
            # bf16 operands move pre-activations by ~1e-2: units within that distance of zero may sit on the other side
            # of the ReLU; the gradient check uses the mask the kernels actually applied (saved activations)
            h = pre * (ctx.hb[0][e, :, :feat] > 0).cpu().double()
        else:
            h = torch.relu(pre)
        y = F.linear(h, p["__head.weight"], p["__head.bias"])
        _close(out[e].cpu().double(), y.detach(), rtol=tol, msg=f"head forward member {e}")
        (y * d_out[e].cpu().double()).sum().backward()
        for k, v in p.items():
            got = net.arena.view(k, e, "grads").cpu().double()
            scale = max(float(v.grad.abs().max()), 1e-6)
            assert float((got - v.grad).abs().max()) <= tol * scale + 1e-7, (precision, k, e)


@pytest.mark.gpu
def test_update_prologue_equals_begin_step_noise_fill_to_bf16():
    """One launch == begin_step + noise_fill + to_bf16 (csrc/cql_fused.cu): same counters, zeroed slots, the same
    Philox stream bit for bit (epoch read before the bump), the same bf16 rows; repeated launches advance the epoch."""
    from d3rlpy_b200._lib import lib

    L = lib()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(0)
    rows, cols, ld = 70, 17, 24
    src = torch.randn(rows, cols, generator=g).to(dev)
    n_norm, n_uni, seed = 70001, 3333, 0x1234ABCD5678
    mask = 0b100101
    for rep in range(2):
        ca = torch.tensor([5, 7, 9, 11, 13, 15], dtype=torch.int32, device=dev) + rep
        cb = ca.clone()
        slots_a, slots_b = torch.ones(64, device=dev), torch.ones(64, device=dev)
        na, nb = torch.zeros(n_norm + n_uni + 3, device=dev), torch.zeros(n_norm + n_uni + 3, device=dev)
        xa = torch.zeros(rows, ld, dtype=torch.bfloat16, device=dev)
        xb = torch.zeros(rows, ld, dtype=torch.bfloat16, device=dev)
        done = torch.zeros(4, dtype=torch.int32, device=dev)
        L.update_prologue(ca.data_ptr(), 6, mask, 0, slots_a.data_ptr(), 64, na.data_ptr(), n_norm, n_uni, seed,
                          src.data_ptr(), cols, rows, cols, xa.data_ptr(), ld, done.data_ptr(), st)
        L.begin_step(cb.data_ptr(), 6, mask, slots_b.data_ptr(), 64, st)
        L.noise_fill(nb.data_ptr(), n_norm, n_uni, seed, cb.data_ptr(), st)
        L.to_bf16(src.data_ptr(), cols, rows, cols, xb.data_ptr(), ld, None, 0, st)
        torch.cuda.synchronize()
        assert torch.equal(ca, cb) and int(done[0]) == 0
        assert torch.equal(slots_a, slots_b) and float(slots_a.abs().sum()) == 0.0
        assert torch.equal(na, nb), "noise stream differs"
        assert torch.equal(xa, xb)
    # no noise, no conversion: counters and slots only
    ca = torch.zeros(6, dtype=torch.int32, device=dev)
    L.update_prologue(ca.data_ptr(), 6, 0b11, 0, slots_a.data_ptr(), 64, None, 0, 0, 0, None, 0, 0, 0, None, 0,
                      done.data_ptr(), st)
    torch.cuda.synchronize()
    assert ca.tolist() == [1, 1, 0, 0, 0, 0]


@pytest.mark.parametrize("nbytes", [1, 15, 16, 256, 4 * 11008, 44032 + 3, 1 << 20])
def test_copy_mapped_round_trip_is_bit_exact(nbytes):
    """csrc/util.cu copy_mapped: pinned host -> device -> pinned host through the kernel (both directions, aligned and
    unaligned ends), bit-exact; a pageable host pointer is refused with an error instead of faulting."""
    from d3rlpy_b200._lib import D3BError, lib

    L, dev, st = lib(), _dev(), _st()
    rs = np.random.RandomState(nbytes % 1000)
    src = torch.from_numpy(rs.randint(0, 256, nbytes + 32).astype(np.uint8)).pin_memory()
    back = torch.zeros(nbytes + 32, dtype=torch.uint8).pin_memory()
    d = torch.zeros(nbytes + 32, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for off in (0, 1):   # 16-byte aligned and misaligned
        d.zero_()
        back.zero_()
        torch.cuda.synchronize()
        L.copy_mapped(d.data_ptr() + off, src.data_ptr() + off, nbytes, st)
        L.copy_mapped(back.data_ptr() + off, d.data_ptr() + off, nbytes, st)
        torch.cuda.synchronize()
        assert torch.equal(d.cpu()[off:off + nbytes], src[off:off + nbytes])
        assert torch.equal(back[off:off + nbytes], src[off:off + nbytes])
        assert int(back[off + nbytes:].sum()) == 0 and int(back[:off].sum()) == 0   # nothing written past the ends
    pageable = torch.zeros(64, dtype=torch.uint8)
    with pytest.raises(D3BError):
        L.copy_mapped(d.data_ptr(), pageable.data_ptr(), 64, st)
    torch.cuda.synchronize()
