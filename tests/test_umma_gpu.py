"""GPU: the tcgen05/TMA/TMEM GEMM (bf16 operands, fp32 accumulate) and the bf16 support kernels vs
torch on the same bf16-rounded operands.  Tolerance: accumulation-order only for fp32 outputs
(1e-4 relative to the row scale), one bf16 ulp (2^-8) for bf16 outputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _st():
    return torch.cuda.current_stream().cuda_stream


def _bf(t):
    return t.to(torch.bfloat16)


def _gemm(L, a, b, m, n, k, E, *, lda, sa, ldb, sb, splits=1, bias=None, relu=0, mask=None, out_bf16=None,
          out_t=None, out_f32=None, atomic=0):
    p = lambda t: None if t is None else t.data_ptr()
    L.umma_gemm(p(a), lda, sa, p(b), ldb, sb, m, n, k, E, splits, p(bias), 0 if bias is None else bias.shape[-1], relu,
                p(mask), 0 if mask is None else mask.shape[-1], 0 if mask is None else mask.shape[-1] * mask.shape[-2],
                p(out_bf16), 0 if out_bf16 is None else out_bf16.shape[-1],
                0 if out_bf16 is None else out_bf16.shape[-1] * out_bf16.shape[-2],
                p(out_t), 0 if out_t is None else out_t.shape[-1], 0 if out_t is None else out_t.shape[-1] * out_t.shape[-2],
                p(out_f32), 0 if out_f32 is None else out_f32.shape[-1],
                0 if out_f32 is None else out_f32.shape[-1] * out_f32.shape[-2], atomic, _st())
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K,E", [(300, 256, 256, 2), (128, 64, 64, 1), (7936, 256, 256, 2), (1000, 400, 320, 1),
                                     # long reductions over few tiles: 128 x 64 tiles, cluster split-K (pixel fc layer)
                                     (256, 750, 752, 1), (32, 512, 3136, 1), (200, 300, 2048, 2), (130, 70, 1032, 3)])
def test_umma_forward_bias_relu_bf16_and_transposed(M, N, K, E):
    from d3rlpy_b200._lib import lib

    L, dev = lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + N)
    a = _bf(torch.randn(E, M, K, generator=g)).to(dev)
    b = _bf(torch.randn(E, N, K, generator=g) / math.sqrt(K)).to(dev)
    bias = torch.randn(E, N, generator=g).to(dev)
    Mp = (M + 7) // 8 * 8
    out = torch.zeros(E, M, N, dtype=torch.bfloat16, device=dev)
    out_t = torch.zeros(E, N, Mp, dtype=torch.bfloat16, device=dev)
    out_f = torch.zeros(E, M, N, device=dev)
    _gemm(L, a, b, M, N, K, E, lda=K, sa=M * K, ldb=K, sb=N * K, bias=bias, relu=1, out_bf16=out, out_t=out_t)
    _gemm(L, a, b, M, N, K, E, lda=K, sa=M * K, ldb=K, sb=N * K, bias=bias, relu=1, out_f32=out_f)
    ref = torch.relu(torch.einsum("emk,enk->emn", a.float(), b.float()) + bias[:, None, :])
    err = (out_f - ref).abs().max().item()
    assert err <= 1e-4 * max(1.0, ref.abs().max().item()), err
    assert torch.equal(out, _bf(out_f))
    assert torch.equal(out_t[:, :, :M], out.transpose(1, 2))


@pytest.mark.parametrize("M,N,K,E", [(256, 400, 1200, 2), (200, 750, 3000, 1), (130, 70, 1030, 3)])
def test_umma_cluster_split_k_with_mask_matches_single_cta(M, N, K, E):
    """dgrad-style launch (ReLU mask, bf16 + fp32 outputs) in the cluster split-K configuration against the same launch
    with the configuration switched off: same values up to the fp32 summation order."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 3 + K)
    ldk, ldn = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    a = torch.full((E, M, ldk), 9.0, dtype=torch.bfloat16)
    a[:, :, :K] = _bf(torch.randn(E, M, K, generator=g))
    b = torch.full((E, N, ldk), -9.0, dtype=torch.bfloat16)
    b[:, :, :K] = _bf(torch.randn(E, N, K, generator=g) / math.sqrt(K))
    a, b = a.to(dev), b.to(dev)
    mask = _bf(torch.randn(E, M, ldn, generator=g)).to(dev)
    outs = {}
    for cluster in (1, 0):
        L.umma_set_cluster(cluster)
        try:
            out = torch.full((E, M, ldn), 3.0, dtype=torch.bfloat16, device=dev)
            out_f = torch.zeros(E, M, N, device=dev)
            _gemm(L, a, b, M, N, K, E, lda=ldk, sa=M * ldk, ldb=ldk, sb=N * ldk, mask=mask, out_bf16=out)
            _gemm(L, a, b, M, N, K, E, lda=ldk, sa=M * ldk, ldb=ldk, sb=N * ldk, mask=mask, out_f32=out_f)
        finally:
            L.umma_set_cluster(1)
        outs[cluster] = (out, out_f)
        assert bool((out[:, :, N:] == 3.0).all()), "padding columns written"
        assert torch.equal(out[:, :, :N], _bf(out_f))
    ref = torch.einsum("emk,enk->emn", a[:, :, :K].float(), b[:, :, :K].float()) * (mask[:, :, :N].float() > 0)
    scale = max(1.0, ref.abs().max().item())
    for cluster in (1, 0):
        assert (outs[cluster][1] - ref).abs().max().item() <= 1e-4 * scale
    assert (outs[1][1] - outs[0][1]).abs().max().item() <= 2e-5 * scale


def test_umma_small_k_zero_fill_shared_a_and_mask():
    """First layer: K=23 (ld 24, TMA zero-fills to 64), input shared by all members; dgrad-style ReLU mask."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), torch.device("cuda:0")
    M, N, K, E, ld = 520, 256, 23, 3, 24
    g = torch.Generator().manual_seed(3)
    a = torch.zeros(M, ld, dtype=torch.bfloat16)
    a[:, :K] = _bf(torch.randn(M, K, generator=g))
    a[:, K:] = 7.0  # padding garbage must be ignored: the tensor map's extent is K, not ld
    b = torch.zeros(E, N, ld, dtype=torch.bfloat16)
    b[:, :, :K] = _bf(torch.randn(E, N, K, generator=g))
    b[:, :, K:] = -3.0
    a, b = a.to(dev), b.to(dev)
    mask = _bf(torch.randn(E, M, N, generator=g)).to(dev)
    out = torch.zeros(E, M, N, dtype=torch.bfloat16, device=dev)
    out_f = torch.zeros(E, M, N, device=dev)
    _gemm(L, a, b, M, N, K, E, lda=ld, sa=0, ldb=ld, sb=N * ld, mask=mask, out_bf16=out)
    _gemm(L, a, b, M, N, K, E, lda=ld, sa=0, ldb=ld, sb=N * ld, mask=mask, out_f32=out_f)
    ref = torch.einsum("mk,enk->emn", a[:, :K].float(), b[:, :, :K].float()) * (mask.float() > 0)
    assert (out_f - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    assert torch.equal(out, _bf(out_f))


@pytest.mark.parametrize("R,No,Ki,E", [(7936, 256, 256, 2), (7936, 256, 23, 2), (256, 256, 17, 1), (200, 40, 24, 1)])
def test_umma_wgrad_split_k_red_accumulate(R, No, Ki, E):
    """dW[No][Ki] += dZ^T[No][R] . (X^T[Ki][R])^T over row splits, RED.ADD into a pre-loaded fp32 buffer."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(R + Ki)
    Rp = (R + 7) // 8 * 8
    dzt = torch.zeros(E, No, Rp, dtype=torch.bfloat16)
    dzt[:, :, :R] = _bf(torch.randn(E, No, R, generator=g) * 0.1)
    xt = torch.zeros(E, Ki, Rp, dtype=torch.bfloat16)
    xt[:, :, :R] = _bf(torch.randn(E, Ki, R, generator=g))
    dzt, xt = dzt.to(dev), xt.to(dev)
    dw0 = torch.randn(E, No, Ki, generator=g).to(dev)
    dw = dw0.clone()
    _gemm(L, dzt, xt, No, Ki, R, E, lda=Rp, sa=No * Rp, ldb=Rp, sb=Ki * Rp, splits=37, out_f32=dw, atomic=1)
    ref = dw0 + torch.einsum("enr,ekr->enk", dzt[:, :, :R].float(), xt[:, :, :R].float())
    assert (dw - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


def test_umma_narrow_n_fp32_store():
    """dgrad restricted to the 6 action columns of layer 1 (B operand = rows [O, O+A) of W1^T)."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), torch.device("cuda:0")
    M, H, O, A, E = 256, 256, 17, 6, 2
    g = torch.Generator().manual_seed(9)
    dz = _bf(torch.randn(E, M, H, generator=g)).to(dev)
    w1t = _bf(torch.randn(E, O + A, H, generator=g)).to(dev)   # [K_in][N_out] = W1^T
    dx = torch.full((E, M, A), 5.0, device=dev)
    bview = w1t[:, O:, :]
    L.umma_gemm(dz.data_ptr(), H, M * H, bview.data_ptr(), H, (O + A) * H, M, A, H, E, 1, None, 0, 0, None, 0, 0, None,
                0, 0, None, 0, 0, dx.data_ptr(), A, M * A, 0, _st())
    torch.cuda.synchronize()
    ref = torch.einsum("emh,eah->ema", dz.float(), w1t[:, O:, :].float())
    assert (dx - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())


def test_bf16_support_kernels():
    from d3rlpy_b200._lib import lib

    L, dev = lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    # to_bf16 with transposed copy
    R, C = 523, 23
    src = torch.randn(R, C, generator=g).to(dev)
    dst = torch.zeros(R, 24, dtype=torch.bfloat16, device=dev)
    dst_t = torch.zeros(C, 528, dtype=torch.bfloat16, device=dev)
    L.to_bf16(src.data_ptr(), C, R, C, dst.data_ptr(), 24, dst_t.data_ptr(), 528, _st())
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :C], _bf(src)) and torch.equal(dst_t[:, :R], _bf(src).t())
    # colsum
    E, M, N = 2, 1000, 256
    dz = _bf(torch.randn(E, M, N, generator=g)).to(dev)
    db = torch.zeros(E, N, device=dev)
    L.colsum_bf16(dz.data_ptr(), N, M * N, db.data_ptr(), N, M, N, E, _st())
    torch.cuda.synchronize()
    ref = dz.float().sum(1)
    assert (db - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    # shadow table: one 5x7 matrix -> row-major ld 8 and transposed ld 8, two members
    msize = 64
    params = torch.randn(2 * msize, generator=g).to(dev)
    shadow = torch.zeros(2 * 128, dtype=torch.bfloat16, device=dev)
    table = torch.tensor([[3 + 1, 5, 7, 0, 8, 40, 8]], dtype=torch.int64)  # src_off 4
    L.shadow_weights(params.data_ptr(), msize, shadow.data_ptr(), 128, table.data_ptr(), 1, 2, _st())
    torch.cuda.synchronize()
    for e in range(2):
        w = params[e * msize + 4:e * msize + 4 + 35].view(5, 7)
        assert torch.equal(shadow[e * 128:e * 128 + 40].view(5, 8)[:, :7], _bf(w))
        assert torch.equal(shadow[e * 128 + 40:e * 128 + 96].view(7, 8)[:, :5], _bf(w).t())
    # heads over bf16 activations
    M, K, N, E = 300, 256, 12, 2
    x = _bf(torch.relu(torch.randn(E, M, K, generator=g))).to(dev)
    w = (torch.randn(E, N, K, generator=g) / 16).to(dev)
    b = torch.randn(E, N, generator=g).to(dev)
    y = torch.zeros(E, M, N, device=dev)
    L.head_forward_bf16(x.data_ptr(), K, M * K, w.data_ptr(), K, N * K, b.data_ptr(), N, y.data_ptr(), N, M * N, M, N,
                        K, E, 0, _st())
    ref = torch.einsum("emk,enk->emn", x.float(), w) + b[:, None, :]
    torch.cuda.synchronize()
    assert (y - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    dy = torch.randn(E, M, N, generator=g).to(dev)
    Mp = (M + 7) // 8 * 8
    dx = torch.zeros(E, M, K, dtype=torch.bfloat16, device=dev)
    dxt = torch.zeros(E, K, Mp, dtype=torch.bfloat16, device=dev)
    L.head_backward_data_bf16(dy.data_ptr(), N, M * N, w.data_ptr(), K, N * K, dx.data_ptr(), K, M * K,
                              dxt.data_ptr(), Mp, K * Mp, x.data_ptr(), K, M * K, M, N, K, E, _st())
    torch.cuda.synchronize()
    ref = torch.einsum("emn,enk->emk", dy, w) * (x.float() > 0)
    assert (dx.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    assert torch.equal(dxt[:, :, :M], dx.transpose(1, 2))
    dw = torch.zeros(E, N, K, device=dev)
    dbias = torch.zeros(E, N, device=dev)
    L.head_backward_weight_bf16(dy.data_ptr(), N, M * N, x.data_ptr(), K, M * K, dw.data_ptr(), K, N * K,
                                dbias.data_ptr(), N, M, N, K, E, _st())
    torch.cuda.synchronize()
    ref = torch.einsum("emn,emk->enk", dy, x.float())
    assert (dw - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    assert (dbias - dy.sum(1)).abs().max().item() <= 1e-4 * dy.sum(1).abs().max().item()


@pytest.mark.parametrize("R,No,Ki,E,shared_b", [(7936, 256, 256, 2, False), (7936, 256, 23, 2, True),
                                                 (256, 256, 17, 1, False), (200, 48, 24, 1, False),
                                                 (1000, 32, 32, 3, False), (130, 256, 119, 2, True)])
def test_umma_wgrad_mn_major_operands(R, No, Ki, E, shared_b):
    """dW[No][Ki] += dZ[R][No]^T . X[R][Ki] with both operands row-major (MN-major UMMA tiles, no transposes)."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(R + Ki)
    ldn, ldk = (No + 7) // 8 * 8, (Ki + 7) // 8 * 8
    dz = torch.full((E, R, ldn), 5.0, dtype=torch.bfloat16)      # padding garbage must be ignored
    dz[:, :, :No] = _bf(torch.randn(E, R, No, generator=g) * 0.1)
    Eb = 1 if shared_b else E
    x = torch.full((Eb, R, ldk), -2.0, dtype=torch.bfloat16)
    x[:, :, :Ki] = _bf(torch.randn(Eb, R, Ki, generator=g))
    dz, x = dz.to(dev), x.to(dev)
    dw0 = torch.randn(E, No, Ki, generator=g).to(dev)
    for splits in (1, 37):
        dw = dw0.clone()
        L.umma_gemm_tn(dz.data_ptr(), ldn, R * ldn, x.data_ptr(), ldk, 0 if shared_b else R * ldk, No, Ki, R, E, splits,
                       dw.data_ptr(), Ki, No * Ki, 1, _st())
        torch.cuda.synchronize()
        ref = dw0 + torch.einsum("ern,erk->enk", dz[:, :, :No].float(), x[:, :, :Ki].float().expand(E, R, Ki))
        err = (dw - ref).abs().max().item()
        assert err <= 2e-4 * max(1.0, ref.abs().max().item()), (splits, err)


@pytest.mark.parametrize("rows,E,in_dim,hidden,n_head,tanh,shared", [
    (7936, 2, 23, [256, 256, 256], 1, False, True),     # c2 critic on the importance-sampling rows
    (512, 1, 17, [256, 256, 256], 12, False, True),     # c2 policy (mu|logstd)
    (300, 3, 14, [256, 256], 1, False, True),           # c1 critic, ragged last tile
    (256, 1, 11, [256, 256], 3, True, True),            # c1 actor with tanh head
    (100, 2, 9, [32, 32, 32], 1, False, True),          # golden-test widths
    (1000, 2, 119, [64, 128, 16, 256], 24, False, True),  # 4 layers, mixed widths, wide head
    (200, 2, 40, [64], 0, False, False),                # single layer, trunk only, per-member input
    (40000, 2, 23, [256, 256, 256], 1, False, True),    # several units per CTA (persistent loop, phase wrap)
    (1, 1, 1, [16], 1, False, True),                    # smallest legal problem
    (129, 3, 256, [240, 48, 80], 5, True, True),        # widths that are not multiples of 64, K_0 = 256
    (127, 2, 7, [16, 256, 16], 32, False, False),       # 32 heads (CUDA-core head, 3-stage ring)
    (513, 1, 65, [192, 192, 192, 192], 2, False, True), # K_0 just above one K block, 4 layers
    (20000, 1, 40, [240, 48, 80, 16], 1, True, True),   # two tiles per unit (157 tiles > SMs), ragged widths, 4 layers
    (19300, 2, 33, [64, 128], 0, False, False),         # two tiles per unit, odd tile count, trunk only, per-member input
    (19000, 1, 23, [256, 256, 256], 1, False, True),    # 149 tiles: last unit of the member has a single tile
])
def test_fused_mlp_forward_matches_torch(rows, E, in_dim, hidden, n_head, tanh, shared):
    """csrc/mlp_fused.cu vs torch on the same bf16-rounded operands (fp32 accumulate, bf16 activations):
    saved activations within one bf16 ulp of the reference chain, head within 2e-3 of its scale."""
    _fused_forward_case(rows, E, in_dim, hidden, n_head, tanh, shared, 0)


@pytest.mark.parametrize("rows,E,save_rows", [(40000, 3, 20000), (50000, 10, 128 * 7), (15872, 2, 7936)])
def test_fused_mlp_forward_partial_save(rows, E, save_rows):
    """`save_rows`: only the tiles below it store their activations (the alpha-step rows of CQL are forward-only).
    Persistent CTAs then alternate between storing and non-storing units (several members, many tiles): the heads of
    ALL rows and the activations of the saved rows must still be right, the other activation rows untouched."""
    _fused_forward_case(rows, E, 23, [256, 256, 256], 1, False, True, save_rows)


def _fused_forward_case(rows, E, in_dim, hidden, n_head, tanh, shared, save_rows):
    import ctypes

    from d3rlpy_b200._lib import lib

    L, dev = lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(rows + in_dim)
    a8 = lambda v: (v + 7) // 8 * 8
    Ex = 1 if shared else E
    x = torch.full((Ex, rows, a8(in_dim)), 3.0, dtype=torch.bfloat16)   # padding garbage must be ignored
    x[:, :, :in_dim] = _bf(torch.randn(Ex, rows, in_dim, generator=g))
    dims = [in_dim] + hidden
    # one flat bf16 shadow and one flat fp32 arena per member, like ParamArena / DenseNet._build_shadow
    w_off, off = [], 0
    for k, n in zip(dims[:-1], dims[1:]):
        w_off.append(off)
        off += n * a8(k)
    sms = a8(off)
    shadow = torch.full((E, sms), -1.0, dtype=torch.bfloat16)
    ws = []
    for o, k, n in zip(w_off, dims[:-1], dims[1:]):
        w = _bf(torch.randn(E, n, k, generator=g) / math.sqrt(k))
        ws.append(w)
        shadow[:, o:o + n * a8(k)].view(E, n, a8(k))[:, :, :k] = w
    feat = hidden[-1]
    nh = max(n_head, 1)
    b_off, off = [], 0
    for n in hidden:
        b_off.append(off)
        off += n
    hw_off, hb_off = off, off + nh * feat
    ms = hb_off + nh
    arena = torch.zeros(E, ms)
    bs = [torch.randn(E, n, generator=g) * 0.1 for n in hidden]
    for o, b in zip(b_off, bs):
        arena[:, o:o + b.shape[1]] = b
    hw = torch.randn(E, nh, feat, generator=g) / math.sqrt(feat)
    hb = torch.randn(E, nh, generator=g) * 0.1
    arena[:, hw_off:hw_off + nh * feat] = hw.reshape(E, -1)
    arena[:, hb_off:hb_off + nh] = hb
    x, shadow, arena = x.to(dev), shadow.to(dev), arena.to(dev)
    acts = [torch.zeros(E, rows, a8(n), dtype=torch.bfloat16, device=dev) for n in hidden]
    out = torch.zeros(E, rows, nh, device=dev)
    nl = len(hidden)
    arr = lambda T, vals: (T * len(vals))(*vals)
    L.mlp_forward_bf16(x.data_ptr(), a8(in_dim), 0 if shared else rows * a8(in_dim), rows, E, nl,
                       arr(ctypes.c_int, dims), arr(ctypes.c_void_p, [shadow.data_ptr() + 2 * o for o in w_off]),
                       arr(ctypes.c_int64, [a8(k) for k in dims[:-1]]), sms,
                       arr(ctypes.c_void_p, [arena.data_ptr() + 4 * o for o in b_off]), ms,
                       arr(ctypes.c_void_p, [a.data_ptr() for a in acts]), arr(ctypes.c_int64, [a.shape[2] for a in acts]),
                       arr(ctypes.c_int64, [a.shape[1] * a.shape[2] for a in acts]),
                       arena.data_ptr() + 4 * hw_off, arena.data_ptr() + 4 * hb_off, ms, n_head, 1 if tanh else 0,
                       out.data_ptr(), save_rows, _st())
    torch.cuda.synchronize()
    saved = rows if not save_rows else min(rows, -(-save_rows // 128) * 128)  # storing is decided per 128-row tile
    h = x[:, :, :in_dim].float().expand(E, rows, in_dim)
    for l, (w, b) in enumerate(zip(ws, bs)):
        ref = torch.relu(torch.einsum("erk,enk->ern", h, w.to(dev).float()) + b.to(dev)[:, None, :])
        got = acts[l][:, :, :hidden[l]].float()
        scale = max(1.0, ref.abs().max().item())
        assert (got[:, :saved] - ref[:, :saved]).abs().max().item() <= 2.0 ** -7 * scale, (l, "saved rows")
        assert saved == rows or got[:, saved:].abs().max().item() == 0.0, (l, "rows that must not be stored")
        # continue from the kernel's own bf16 activations so that errors do not compound in the check
        h = got if saved == rows else torch.cat([got[:, :saved], ref[:, saved:].to(torch.bfloat16).float()], 1)
    if n_head:
        ref = torch.einsum("erk,ejk->erj", h, hw.to(dev)) + hb.to(dev)[:, None, :]
        if tanh:
            ref = torch.tanh(ref)
        err = (out - ref).abs().max().item()
        assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("rows,E,in_dim,hidden,n_head,wg,dxr", [
    (7936, 2, 23, [256, 256, 256], 1, True, None),       # c2 critic step
    (256, 2, 23, [256, 256, 256], 1, False, (17, 6)),    # c2 actor step: dX of the action columns only
    (256, 1, 17, [256, 256, 256], 12, True, None),       # c2 policy
    (300, 3, 14, [256, 256], 1, True, (11, 3)),          # ragged tile, weight grads + dx
    (100, 2, 9, [32, 32, 32], 1, True, None),            # golden-test widths
    (1000, 2, 119, [64, 128, 16, 256], 8, True, (100, 19)),
    (200, 2, 40, [64], 3, True, (0, 40)),                # single layer
    (40000, 2, 23, [256, 256, 256], 1, True, None),      # several units per CTA
    (1, 1, 1, [16], 1, True, (0, 1)),                    # smallest legal problem
    (129, 3, 256, [240, 48, 80], 5, True, (250, 6)),     # widths that are not multiples of 64, dx at the right edge
    (127, 2, 9, [16, 256, 16], 16, True, (1, 8)),        # 16 heads
    (513, 1, 65, [192, 192, 192, 192], 2, False, (3, 61)),
])
def test_fused_mlp_backward_matches_torch(rows, E, in_dim, hidden, n_head, wg, dxr):
    """csrc/mlp_fused.cu backward chain vs torch on the same bf16 operands, stage by stage."""
    import ctypes

    from d3rlpy_b200._lib import lib

    L, dev = lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(rows + in_dim + 1)
    a8 = lambda v: (v + 7) // 8 * 8
    dims = [in_dim] + hidden
    nl = len(hidden)
    w_off, off = [], 0
    for k, n in zip(dims[:-1], dims[1:]):
        w_off.append(off)
        off += n * a8(k)
    sms = a8(off)
    shadow = torch.full((E, sms), -1.0, dtype=torch.bfloat16)
    ws = []
    for o, k, n in zip(w_off, dims[:-1], dims[1:]):
        w = _bf(torch.randn(E, n, k, generator=g) / math.sqrt(k))
        ws.append(w.to(dev))
        shadow[:, o:o + n * a8(k)].view(E, n, a8(k))[:, :, :k] = w
    shadow = shadow.to(dev)
    feat = hidden[-1]
    # saved activations: random bf16 with ~half of the entries <= 0 (the ReLU mask only looks at the sign)
    acts = []
    for n in hidden:
        a = torch.zeros(E, rows, a8(n), dtype=torch.bfloat16)
        a[:, :, :n] = _bf(torch.relu(torch.randn(E, rows, n, generator=g)))
        acts.append(a.to(dev))
    # gradient arena layout: [b_0 .. b_{L-1} | head_w | head_b]
    b_off, off = [], 0
    for n in hidden:
        b_off.append(off)
        off += n
    hw_off, hb_off = off, off + n_head * feat
    ms = (hb_off + n_head + 3) // 4 * 4
    params = torch.zeros(E, ms)
    hw = torch.randn(E, n_head, feat, generator=g) / math.sqrt(feat)
    params[:, hw_off:hw_off + n_head * feat] = hw.reshape(E, -1)
    params = params.to(dev)
    hw = hw.to(dev)
    grads = torch.zeros(E, ms, device=dev)
    d_head = (torch.randn(E, rows, n_head, generator=g) / rows).to(dev)
    dz = [torch.zeros(E, rows, a8(n), dtype=torch.bfloat16, device=dev) for n in hidden]
    dx = torch.zeros(E, rows, dxr[1], device=dev) if dxr else None
    arr = lambda T, vals: (T * len(vals))(*vals)
    L.mlp_backward_bf16(rows, E, nl, arr(ctypes.c_int, dims),
                        arr(ctypes.c_void_p, [shadow.data_ptr() + 2 * o for o in w_off]),
                        arr(ctypes.c_int64, [a8(k) for k in dims[:-1]]), sms,
                        arr(ctypes.c_void_p, [a.data_ptr() for a in acts]), arr(ctypes.c_int64, [a.shape[2] for a in acts]),
                        arr(ctypes.c_int64, [a.shape[1] * a.shape[2] for a in acts]),
                        arr(ctypes.c_void_p, [d.data_ptr() for d in dz]) if wg else None,
                        arr(ctypes.c_int64, [d.shape[2] for d in dz]) if wg else None,
                        arr(ctypes.c_int64, [d.shape[1] * d.shape[2] for d in dz]) if wg else None,
                        d_head.data_ptr(), params.data_ptr() + 4 * hw_off, ms, n_head,
                        arr(ctypes.c_void_p, [grads.data_ptr() + 4 * o for o in b_off]) if wg else None,
                        grads.data_ptr() + 4 * hw_off if wg else None, grads.data_ptr() + 4 * hb_off if wg else None, ms,
                        None, dx.data_ptr() if dxr else None, dxr[1] if dxr else 0, rows * dxr[1] if dxr else 0,
                        dxr[0] if dxr else 0, dxr[1] if dxr else 0, _st())
    torch.cuda.synchronize()

    def close(got, ref, tol, what):
        err = (got - ref).abs().max().item()
        assert err <= tol * max(ref.abs().max().item(), 1e-30), (what, err, ref.abs().max().item())

    # reference chain in fp32 over the same bf16 operands
    cur = (torch.einsum("erj,ejc->erc", d_head, hw) * (acts[-1][:, :, :feat].float() > 0))
    for l in range(nl - 1, -1, -1):
        n = hidden[l]
        cur_b = _bf(cur).float()
        if wg:
            close(dz[l][:, :, :n].float(), cur_b, 2.0 ** -7, f"dz{l}")
            cur_b = dz[l][:, :, :n].float()  # continue from the kernel's own rounding
            close(grads[:, b_off[l]:b_off[l] + n], cur_b.sum(1), 2e-3, f"dbias{l}")
        if l > 0:
            cur = torch.einsum("ern,enk->erk", cur_b, ws[l].float()) * (acts[l - 1][:, :, :hidden[l - 1]].float() > 0)
        elif dxr:
            ref = torch.einsum("ern,enk->erk", cur_b, ws[0].float()[:, :, dxr[0]:dxr[0] + dxr[1]])
            close(dx, ref, 2.0 ** -6 if not wg else 2e-3, "dx")
    if wg:
        close(grads[:, hw_off:hw_off + n_head * feat].view(E, n_head, feat),
              torch.einsum("erj,erc->ejc", d_head, acts[-1][:, :, :feat].float()), 2e-3, "d_head_w")
        close(grads[:, hb_off:hb_off + n_head], d_head.sum(1), 2e-3, "d_head_b")


@pytest.mark.parametrize("M,K,N,E,tanh", [(300, 750, 6, 1, True), (1000, 300, 1, 2, False), (130, 1000, 32, 1, False),
                                          (64, 516, 7, 3, False)])
def test_head_forward_bf16_wide_inputs(M, K, N, E, tanh):
    """Heads over wide bf16 activations (BCQ's 750 / 300 features): the 16-byte-chunk kernel, incl. K not a multiple of
    8 (tail elements), padded leading dimension with garbage, tanh."""
    from d3rlpy_b200._lib import lib

    L, dev = lib(), torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + K)
    ld = (K + 7) // 8 * 8
    x = torch.full((E, M, ld), 77.0, dtype=torch.bfloat16)
    x[:, :, :K] = _bf(torch.relu(torch.randn(E, M, K, generator=g)))
    x = x.to(dev)
    w = (torch.randn(E, N, K, generator=g) / math.sqrt(K)).to(dev)
    b = torch.randn(E, N, generator=g).to(dev)
    y = torch.zeros(E, M, N, device=dev)
    L.head_forward_bf16(x.data_ptr(), ld, M * ld, w.data_ptr(), K, N * K, b.data_ptr(), N, y.data_ptr(), N, M * N, M, N,
                        K, E, 1 if tanh else 0, _st())
    torch.cuda.synchronize()
    ref = torch.einsum("emk,enk->emn", x[:, :, :K].float(), w) + b[:, None, :]
    if tanh:
        ref = torch.tanh(ref)
    assert (y - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
