"""GPU: the 3xTF32 tensor-core GEMM (csrc/gemm_tf32x3.cu, the fp32-mode engine) through the C ABI against an fp64
PyTorch statement of the same contraction.

Tolerance (stated per assertion): every output element within 3e-6 * sum_r |a||b| of the fp64 result — the
error-compensated split keeps 22 significant bits per operand and the remaining error is the tensor core's
round-toward-zero accumulation (profiles/r2/tc32_probe.py), so the bound is that of an fp32 FMA chain of this
length (the SIMT kernel measures 2e-7 ... 1e-6 on the same data; an uncompensated TF32 product is off by ~1e-4 of
that sum and bf16 by ~1e-3: the same test fails for them by a wide margin, asserted at the end)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BOUND = 3e-6


def _dev():
    return torch.device("cuda:0")


def _st():
    return torch.cuda.current_stream().cuda_stream


def _check(got, ref64, bound64, what):
    err = (got.double() - ref64).abs()
    lim = BOUND * bound64 + 1e-30
    worst = float((err / lim).max())
    assert worst <= 1.0, f"{what}: worst err/bound {worst:.3f} (max abs err {float(err.max()):.3e})"


def _rand(g, *shape, scale=1.0):
    # values with full 24-bit mantissas and a spread of magnitudes (a hi-only TF32 product would miss the bound)
    return (torch.randn(*shape, generator=g) * torch.exp(torch.randn(*shape, generator=g)) * scale)


@pytest.mark.parametrize("M,N,K,E,shared,ldpad", [
    (256, 256, 23, 2, True, 0),       # c2 critic layer 0: K = obs + act, unaligned rows -> scalar loads
    (15872, 256, 256, 2, False, 0),   # c2 importance-sampling rows
    (130, 70, 45, 3, False, 3),       # ragged everything, padded leading dimensions
    (512, 300, 400, 1, False, 0),     # BCQ widths (N tiles 128+128+44)
    (64, 32, 256, 1, False, 0),       # half a row tile, the narrowest N tile
    (1000, 128, 512, 5, False, 4),
    (256, 256, 256, 2, False, 0),     # latency configuration: 128 x 64 tiles, cluster split-K over 4 CTAs
    (512, 256, 256, 1, False, 0),
    (200, 100, 300, 2, False, 3),     # ragged tile + ragged K split (19 K blocks over 4 CTAs), padded rows
    (100, 750, 750, 1, False, 0),     # BCQ widths at batch 100
])
def test_tc32_forward(M, N, K, E, shared, ldpad):
    from d3rlpy_b200._lib import lib

    L, dev = lib(), _dev()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    ldx, ldw, ldy = K + ldpad, K + ldpad, N + ldpad
    x = torch.full((1 if shared else E, M, ldx), float("nan"))
    w = torch.full((E, N, ldw), float("nan"))
    x[..., :K] = _rand(g, x.shape[0], M, K)
    w[..., :K] = _rand(g, E, N, K, scale=1.0 / math.sqrt(K))
    b = torch.randn(E, N, generator=g)
    x, w, b = x.to(dev), w.to(dev), b.to(dev)
    y = torch.full((E, M, ldy), -7.0, device=dev)
    L.tc32_gemm(x.data_ptr(), ldx, 0 if shared else M * ldx, 1, w.data_ptr(), ldw, N * ldw, 1, y.data_ptr(), ldy,
                M * ldy, M, N, K, E, 1, b.data_ptr(), N, 1, None, 0, 0, None, 0, 0, 0, _st())
    torch.cuda.synchronize()
    xe = x[..., :K].expand(E, M, K).double()
    we = w[..., :K].double()
    pre = torch.einsum("emk,enk->emn", xe, we) + b.double()[:, None, :]
    bound = torch.einsum("emk,enk->emn", xe.abs(), we.abs()) + b.double().abs()[:, None, :]
    _check(y[..., :N], torch.relu(pre), bound, "forward")
    if ldpad:
        assert bool((y[..., N:] == -7.0).all()), "padding columns of the output were written"


@pytest.mark.parametrize("M,N,K,E,cols", [(15872, 256, 256, 2, None), (256, 256, 23, 2, None), (256, 256, 23, 2, (17, 6)),
                                          (130, 70, 45, 3, None), (512, 300, 400, 1, None),
                                          (256, 256, 256, 2, None), (200, 300, 100, 2, None)])   # cluster split-K
def test_tc32_backward_data(M, N, K, E, cols):
    """dx = (dy W) * [src > 0]; `cols` = (first column, count) restricts to a column range of W (actor step)."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), _dev()
    g = torch.Generator().manual_seed(M + N * 5 + K * 11)
    dy = _rand(g, E, M, N).to(dev)
    w = _rand(g, E, N, K, scale=1.0 / math.sqrt(N)).to(dev)
    c0, nc = cols if cols else (0, K)
    src = torch.randn(E, M, nc, generator=g).to(dev) if not cols else None
    dx = torch.full((E, M, nc), -7.0, device=dev)
    L.tc32_gemm(dy.data_ptr(), N, M * N, 1, w.data_ptr() + 4 * c0, K, N * K, 0, dx.data_ptr(), nc, M * nc, M, nc, N, E,
                1, None, 0, 0, src.data_ptr() if src is not None else None, nc, M * nc, None, 0, 0, 0, _st())
    torch.cuda.synchronize()
    ws = w[:, :, c0:c0 + nc].double()
    ref = torch.einsum("emn,enk->emk", dy.double(), ws)
    bound = torch.einsum("emn,enk->emk", dy.double().abs(), ws.abs())
    if src is not None:
        ref = ref * (src > 0).double()
    _check(dx, ref, bound, "dgrad")


@pytest.mark.parametrize("M,N,K,E,shared,splits", [(15872, 256, 256, 2, False, 18), (256, 256, 23, 2, True, 4),
                                                  (130, 70, 45, 3, False, 1), (512, 300, 400, 1, False, 7),
                                                  (2560, 64, 17, 1, True, 40)])
def test_tc32_backward_weight(M, N, K, E, shared, splits):
    """dW += dy^T x (split over the M minibatch rows, RED epilogue), db += column sums of dy."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), _dev()
    g = torch.Generator().manual_seed(M * 3 + N + K * 13)
    dy = _rand(g, E, M, N).to(dev)
    x = _rand(g, 1 if shared else E, M, K).to(dev)
    dw0 = torch.randn(E, N, K, generator=g).to(dev)
    db0 = torch.randn(E, N, generator=g).to(dev)
    dw, db = dw0.clone(), db0.clone()
    L.tc32_gemm(dy.data_ptr(), N, M * N, 0, x.data_ptr(), K, 0 if shared else M * K, 0, dw.data_ptr(), K, N * K, N, K,
                M, E, splits, None, 0, 0, None, 0, 0, db.data_ptr(), N, 1, _st())
    torch.cuda.synchronize()
    xe = x.expand(E, M, K).double()
    ref = dw0.double() + torch.einsum("emn,emk->enk", dy.double(), xe)
    bound = dw0.double().abs() + torch.einsum("emn,emk->enk", dy.double().abs(), xe.abs())
    # the RED epilogue adds `splits` partial sums in fp32 on top of the in-tile accumulation
    _check(dw, ref, bound * 2, "wgrad")
    refb = db0.double() + dy.double().sum(1)
    boundb = db0.double().abs() + dy.double().abs().sum(1)
    _check(db, refb, boundb * 8, "bias gradient (fp32 column sums)")


def test_tc32_beats_single_tf32_and_matches_simt():
    """The engine switch: linear_forward through both engines on the same operands.  Both meet the fp32 bound; a
    plain (uncompensated) TF32 product, emulated by truncating the operands, does not — so the bound is meaningful."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), _dev()
    g = torch.Generator().manual_seed(3)
    M, N, K, E = 2048, 256, 256, 2
    x, w = _rand(g, E, M, K).to(dev), _rand(g, E, N, K, scale=1 / 16).to(dev)
    b = torch.zeros(E, N, device=dev)
    ref = torch.einsum("emk,enk->emn", x.double(), w.double())
    bound = torch.einsum("emk,enk->emn", x.double().abs(), w.double().abs())
    prev = L.raw("d3b_get_fp32_engine")()
    try:
        for eng in (0, 1):
            L.set_fp32_engine(eng)
            y = torch.empty(E, M, N, device=dev)
            L.linear_forward(x.data_ptr(), K, M * K, w.data_ptr(), K, N * K, b.data_ptr(), N, y.data_ptr(), N, M * N, M, N,
                             K, E, 0, _st())
            torch.cuda.synchronize()
            _check(y, ref, bound, f"linear_forward engine {eng}")
    finally:
        L.set_fp32_engine(prev)
    trunc = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    y1 = torch.einsum("emk,enk->emn", trunc(x).double(), trunc(w).double())
    assert float(((y1 - ref).abs() / (BOUND * bound)).max()) > 20.0


@pytest.mark.parametrize("force_s", [1, 2, 8])
def test_tc32_cluster_split_sizes(force_s):
    """Every cluster size gives the same result within the bound (forced through the profiling variant word)."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), _dev()
    g = torch.Generator().manual_seed(force_s)
    M, N, K, E = 256, 192, 512, 2
    x, w = _rand(g, E, M, K).to(dev), _rand(g, E, N, K, scale=1 / math.sqrt(K)).to(dev)
    b = torch.randn(E, N, generator=g).to(dev)
    y = torch.empty(E, M, N, device=dev)
    L.tc32_set_variant(force_s << 8)
    try:
        L.tc32_gemm(x.data_ptr(), K, M * K, 1, w.data_ptr(), K, N * K, 1, y.data_ptr(), N, M * N, M, N, K, E, 1,
                    b.data_ptr(), N, 1, None, 0, 0, None, 0, 0, 0, _st())
        torch.cuda.synchronize()
    finally:
        L.tc32_set_variant(0)
    pre = torch.einsum("emk,enk->emn", x.double(), w.double()) + b.double()[:, None, :]
    bound = torch.einsum("emk,enk->emn", x.double().abs(), w.double().abs()) + b.double().abs()[:, None, :]
    _check(y, torch.relu(pre), bound, f"forward, cluster of {force_s}")
