"""Dense networks over flat arenas: ReLU MLP trunk (+ narrow head) for all E ensemble members in
one launch per layer.  Mirrors VectorEncoder/VectorEncoderWithAction + `_fc`/`_mu`/`_logstd` heads
(d3rlpy/models/torch/encoders.py:236-339, q_functions/mean_q_function.py:61-72,
policies.py:47-59,82-97,127-181, imitators.py:13-72) without nn.Module or autograd: forward and the
hand-written backward are sequences of C-ABI kernel launches on the caller's stream.

Two arithmetic modes:
  * "fp32": SIMT FFMA GEMMs, everything fp32 (parity tolerance 1e-5).
  * "bf16": tcgen05 tensor-core GEMMs over bf16 K-major shadows of weights/activations with fp32
    accumulation, fp32 master weights, fp32 heads/losses/optimizer (parity tolerance 1e-2).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ._lib import lib
from .arena import ParamArena


def _p(t) -> Optional[int]:
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def _a8(n: int) -> int:
    return (n + 7) // 8 * 8


HEAD_NARROW_MAX = 32  # csrc/heads.cu: warp-per-row GEMV heads; wider heads (quantile heads) run as dense layers


def _wide_head_forward(net, which: str, cur, ld: int, sx: int, rows: int, E: int, head_out, stream: int, member0=0):
    """Head wider than 32 outputs (DiscreteQRQFunction: Linear(feature, A * n_quantiles), qr_q_function.py:33-36):
    one more dense layer without activation and with fp32 output — SIMT GEMM in fp32 mode, tcgen05 GEMM over the bf16
    head-weight shadow in bf16 mode."""
    L, ms, n, d = lib(), net.arena.member_size, net.head_out, net.feat
    hw = net.arena.addr(which, "__head.weight", member0)
    hb = net.arena.addr(which, "__head.bias", member0)
    if net.precision == "fp32":
        L.linear_forward(cur, ld, sx, hw, d, ms, hb, ms, _p(head_out), n, rows * n, rows, n, d, E, 0, stream)
        return
    base = net.shadow if which == "params" else net.shadow_target
    w_off, ldw = net._sh_w[net._head_shadow]
    wptr = base.data_ptr() + 2 * (member0 * net.shadow_member + w_off)
    L.umma_gemm(cur, ld, sx, wptr, ldw, net.shadow_member, rows, n, d, E, 1, hb, ms, 0, None, 0, 0, None, 0, 0,
                None, 0, 0, _p(head_out), n, rows * n, 0, stream)


def _wide_head_backward(net, ctx, rows: int, E: int, d_head, last, ldl: int, sl: int, dcur, ldd: int, sd: int,
                        stream: int, weight_grads: bool = True, member0=0):
    """dW_head += d_head^T H, db_head += colsum(d_head), dZ = (d_head W_head) * [H > 0] for a wide head."""
    L, ms, n, d = lib(), net.arena.member_size, net.head_out, net.feat
    hw_g = net.arena.addr("grads", "__head.weight", member0)
    hb_g = net.arena.addr("grads", "__head.bias", member0)
    if net.precision == "fp32":
        if weight_grads:
            L.linear_backward_weight(_p(d_head), n, rows * n, last, ldl, sl, hw_g, d, ms, hb_g, ms, rows, n, d, E, stream)
        L.linear_backward_data(_p(d_head), n, rows * n, net.arena.addr("params", "__head.weight", member0), d, ms, dcur,
                               ldd, sd, last, ldl, sl, rows, n, d, E, stream)
        return
    ln = _a8(n)
    dh = getattr(ctx, "dh_wide", None)
    if dh is None or dh.shape != (E, rows, ln):
        dh = ctx.dh_wide = torch.zeros(E, rows, ln, dtype=torch.bfloat16, device=net.device)
    L.to_bf16(_p(d_head), n, E * rows, n, _p(dh), ln, None, 0, stream)
    if weight_grads:
        L.colsum_bf16(_p(dh), ln, rows * ln, hb_g, ms, rows, n, E, stream)
        tiles = -(-n // 128) * -(-d // 256) * E
        splits = max(1, min(-(-rows // 64), -(-148 // tiles)))
        L.umma_gemm_tn(_p(dh), ln, rows * ln, last, ldl, sl, n, d, rows, E, splits, hw_g, d, ms, 1, stream)
    wt_off, ldwt = net._sh_wt[net._head_shadow]
    wt = net.shadow.data_ptr() + 2 * (member0 * net.shadow_member + wt_off)
    L.umma_gemm(_p(dh), ln, rows * ln, wt, ldwt, net.shadow_member, rows, d, n, E, 1, None, 0, 0, last, ldl, sl, dcur,
                ldd, sd, None, 0, 0, None, 0, 0, 0, stream)


class Ctx:
    """Per-call-site workspace of one forward(/backward): saved activations and gradient scratch."""

    def __init__(self, net: "DenseNet", rows: int, members: int, train: bool):
        self.rows, self.members, self.train = rows, members, train
        dev, E = net.device, members
        hid = net.hidden
        if net.precision == "fp32":
            self.acts = [torch.zeros(E, rows, h, dtype=torch.float32, device=dev) for h in hid]
            # dZ_l of every layer: the weight-gradient GEMMs run on a side stream beside the data-gradient chain, so a
            # layer's dZ must stay intact while the chain moves on (no ping-pong scratch)
            self.dz = [torch.zeros(E, rows, h, dtype=torch.float32, device=dev) for h in hid] if train else None
        else:
            bf = torch.bfloat16
            self.ldk0 = _a8(net.in_dim)
            self.xb = torch.zeros(rows, self.ldk0, dtype=bf, device=dev)
            # saved activations H_l (row-major bf16): ReLU masks and weight-gradient operands of the backward pass
            self.hb = [torch.zeros(E, rows, _a8(h), dtype=bf, device=dev) for h in hid]
            if train:
                hm = _a8(max(hid + [net.in_dim]))
                # dZ_l for every layer (the weight-gradient GEMMs read them after the data-gradient chain)
                self.dz = [torch.zeros(E, rows, _a8(h), dtype=bf, device=dev) for h in hid]


class DenseNet:
    """trunk: in_dim -> hidden[0] -> ... -> hidden[-1] (ReLU), head: hidden[-1] -> head_out."""

    def __init__(self, in_dim: int, hidden: Sequence[int], heads: Sequence[Tuple[str, int]], members: int,
                 device, trunk_prefix: str, member_key: str = "{name}", with_target: bool = False,
                 seed_gen: Optional[torch.Generator] = None, precision: str = "fp32"):
        assert precision in ("fp32", "bf16")
        self.in_dim, self.hidden, self.members = in_dim, list(hidden), members
        self.heads = list(heads)
        self.head_out = sum(n for _, n in heads)
        self.trunk_prefix = trunk_prefix
        self.precision = precision
        entries = []
        d = in_dim
        for i, h in enumerate(self.hidden):
            entries.append((f"{trunk_prefix}_fcs.{i}.weight", (h, d)))
            entries.append((f"{trunk_prefix}_fcs.{i}.bias", (h,)))
            d = h
        # all heads are ONE allocation (rows concatenated) so they run as one narrow GEMV; the
        # reference's per-head names are exported as row slices of it
        entries.append(("__head.weight", (self.head_out, d)))
        entries.append(("__head.bias", (self.head_out,)))
        # registration order of the reference modules: head by head, weight then bias (it is also the parameter
        # order of the torch optimizers, i.e. the index space of their state_dict)
        exports, r0 = [], 0
        for name, n in self.heads:
            exports.append((f"{name}.weight", "__head.weight", r0, n))
            exports.append((f"{name}.bias", "__head.bias", r0, n))
            r0 += n
        self.arena = ParamArena(entries, members, device, with_target=with_target, member_key=member_key,
                                exports=exports)
        self.feat = d
        self.device = device
        self.wide_head = self.head_out > HEAD_NARROW_MAX
        # whole-network fused forward (csrc/mlp_fused.cu): widths multiples of 16 and <= 256, <= 4 layers, head <= 32
        self.fused_ok = (precision == "bf16" and 1 <= len(self.hidden) <= 4 and in_dim <= 256 and self.head_out <= 32
                         and all(h % 16 == 0 and 16 <= h <= 256 for h in self.hidden))
        self._ctx: Dict[str, Ctx] = {}
        self._init_params(seed_gen)
        if precision == "bf16":
            self._build_shadow(with_target)

    def _init_params(self, gen):
        """nn.Linear default init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias."""
        a = self.arena
        with torch.no_grad():
            for e in range(self.members):
                for name, shape in a.shapes.items():
                    fan_in = shape[1] if len(shape) == 2 else a.shapes[name[:-4] + "weight"][1]
                    bound = 1.0 / math.sqrt(fan_in)
                    v = (torch.rand(shape, generator=gen) * 2 - 1) * bound
                    a.view(name, e).copy_(v)
        a.sync_target_from_params()

    # ------------------------------------------------------------------ bf16 shadows of the trunk weights
    def _build_shadow(self, with_target: bool):
        """Per member: for every trunk layer a row-major bf16 copy W [N][ld8(K)] (forward B operand) and a
        transposed copy W^T [K][ld8(N)] (dgrad B operand) — both K-major for the tcgen05 GEMM."""
        rows, off = [], 0
        self._sh_w, self._sh_wt = [], []
        d = self.in_dim
        for i, h in enumerate(self.hidden):
            ldk, ldn = _a8(d), _a8(h)
            w_off = off
            off += h * ldk
            wt_off = off
            off += d * ldn
            self._sh_w.append((w_off, ldk))
            self._sh_wt.append((wt_off, ldn))
            rows.append([self.arena.offsets[f"{self.trunk_prefix}_fcs.{i}.weight"], h, d, w_off, ldk, wt_off, ldn])
            d = h
        if self.wide_head:  # a wide head is one more tensor-core layer: shadow its weight like the trunk's
            ldk, ldn = _a8(d), _a8(self.head_out)
            self._head_shadow = len(self._sh_w)
            self._sh_w.append((off, ldk))
            self._sh_wt.append((off + self.head_out * ldk, ldn))
            rows.append([self.arena.offsets["__head.weight"], self.head_out, d, off, ldk, off + self.head_out * ldk, ldn])
            off += self.head_out * ldk + d * ldn
        self.shadow_member = _a8(off)
        self._table = torch.tensor(rows, dtype=torch.int64)
        n = self.shadow_member * self.members
        self.shadow = torch.zeros(n, dtype=torch.bfloat16, device=self.device)
        self.shadow_target = torch.zeros(n, dtype=torch.bfloat16, device=self.device) if with_target else None

    def refresh_shadow(self, which: str, stream: int):
        """Re-derive the bf16 shadows from the fp32 master weights (after Adam / soft_sync / load)."""
        if self.precision != "bf16":
            return
        src = self.arena.params if which == "params" else self.arena.target
        dst = self.shadow if which == "params" else self.shadow_target
        lib().shadow_weights(_p(src), self.arena.member_size, _p(dst), self.shadow_member, self._table.data_ptr(),
                             self._table.shape[0], self.members, stream)

    def _sw(self, which, i, member0=0):
        base = self.shadow if which == "params" else self.shadow_target
        return base.data_ptr() + 2 * (member0 * self.shadow_member + self._sh_w[i][0]), self._sh_w[i][1]

    def _swt(self, i, member0=0):
        return self.shadow.data_ptr() + 2 * (member0 * self.shadow_member + self._sh_wt[i][0]), self._sh_wt[i][1]

    # ------------------------------------------------------------------ pointers
    def _w(self, which, i, member=0):
        return self.arena.addr(which, f"{self.trunk_prefix}_fcs.{i}.weight", member)

    def _b(self, which, i, member=0):
        return self.arena.addr(which, f"{self.trunk_prefix}_fcs.{i}.bias", member)

    def _hw(self, which, member=0):
        return self.arena.addr(which, "__head.weight", member)

    def _hb(self, which, member=0):
        return self.arena.addr(which, "__head.bias", member)

    wgrad_side = True   # fp32 mode: weight-gradient GEMMs on a side stream beside the data-gradient chain

    def _wgrad_stream(self) -> int:
        if getattr(self, "_wg_side", None) is None:
            self._wg_side = torch.cuda.Stream(device=self.device)
        return self._wg_side.cuda_stream

    def ctx(self, tag: str, rows: int, members: Optional[int] = None, train: bool = True) -> Ctx:
        E = members or self.members
        c = self._ctx.get(tag)
        if c is None or c.rows != rows or c.members != E or c.train != train:
            c = Ctx(self, rows, E, train)
            self._ctx[tag] = c
        return c

    # ------------------------------------------------------------------ forward
    def forward(self, which: str, x, ldx: int, rows: int, ctx: Ctx, head_out, stream: int,
                head_tanh: bool = False, member0: int = 0, x_bf16=None, save_rows: int = 0):
        """x: fp32 [rows, ldx], shared by all members.  Saves activations in ctx; head_out: fp32
        [E, rows, head_out] (None = trunk only)."""
        L = lib()
        E = ctx.members
        ms = self.arena.member_size
        n = self.head_out
        if self.precision == "fp32":
            cur, ld, sx = _p(x), ldx, 0
            d = self.in_dim
            for i, h in enumerate(self.hidden):
                y = ctx.acts[i]
                # member stride = the workspace's row count: a later backward may cover a prefix of the rows only
                L.linear_forward(cur, ld, sx, self._w(which, i, member0), d, ms, self._b(which, i, member0), ms,
                                 _p(y), h, ctx.rows * h, rows, h, d, E, 1, stream)
                cur, ld, sx, d = _p(y), h, ctx.rows * h, h
            if head_out is not None and self.wide_head:
                assert not head_tanh
                _wide_head_forward(self, which, cur, ld, sx, rows, E, head_out, stream, member0)
            elif head_out is not None:
                L.head_forward(cur, ld, sx, self._hw(which, member0), d, ms, self._hb(which, member0), ms,
                               _p(head_out), n, rows * n, rows, n, d, E, 1 if head_tanh else 0, stream)
            return
        # ---- bf16 mode.  x_bf16 = (device pointer, ld): rows already assembled as bf16 GEMM operands
        if x_bf16 is not None:
            ctx.x_src = (int(x_bf16[0]), int(x_bf16[1]))
        else:
            L.to_bf16(_p(x), ldx, rows, self.in_dim, _p(ctx.xb), ctx.ldk0, None, 0, stream)
            ctx.x_src = (_p(ctx.xb), ctx.ldk0)
        xs, xld = ctx.x_src
        sms = self.shadow_member
        if self.fused_ok:
            # ONE persistent launch: trunk + head for all members, activations chained on-chip
            import ctypes

            nl = len(self.hidden)
            dims = (ctypes.c_int * (nl + 1))(self.in_dim, *self.hidden)
            wp = (ctypes.c_void_p * nl)(*[self._sw(which, i, member0)[0] for i in range(nl)])
            ldw = (ctypes.c_int64 * nl)(*[self._sw(which, i, member0)[1] for i in range(nl)])
            bp = (ctypes.c_void_p * nl)(*[self._b(which, i, member0) for i in range(nl)])
            if ctx.train:
                ap = (ctypes.c_void_p * nl)(*[_p(t) for t in ctx.hb])
                lda = (ctypes.c_int64 * nl)(*[t.shape[2] for t in ctx.hb])
                sa = (ctypes.c_int64 * nl)(*[t.shape[1] * t.shape[2] for t in ctx.hb])
            else:
                ap, lda, sa = None, None, None
            with_head = head_out is not None
            L.mlp_forward_bf16(xs, xld, 0, rows, E, nl, dims, wp, ldw, sms, bp, ms, ap, lda, sa,
                               self._hw(which, member0) if with_head else None,
                               self._hb(which, member0) if with_head else None, ms, n if with_head else 0,
                               1 if head_tanh else 0, _p(head_out) if with_head else None, save_rows, stream)
            return
        cur, ld, sx = xs, xld, 0
        d = self.in_dim
        for i, h in enumerate(self.hidden):
            wptr, ldw = self._sw(which, i, member0)
            y = ctx.hb[i]
            lh = _a8(h)
            L.umma_gemm(cur, ld, sx, wptr, ldw, sms, rows, h, d, E, 1, self._b(which, i, member0), ms, 1,
                        None, 0, 0, _p(y), lh, rows * lh, None, 0, 0, None, 0, 0, 0, stream)
            cur, ld, sx, d = _p(y), lh, rows * lh, h
        if head_out is not None and self.wide_head:
            assert not head_tanh
            _wide_head_forward(self, which, cur, ld, sx, rows, E, head_out, stream, member0)
        elif head_out is not None:
            L.head_forward_bf16(cur, ld, sx, self._hw(which, member0), d, ms, self._hb(which, member0), ms,
                                _p(head_out), n, rows * n, rows, n, d, E, 1 if head_tanh else 0, stream)

    # ------------------------------------------------------------------ backward
    def backward(self, x, ldx: int, rows: int, ctx: Ctx, d_head, stream: int, weight_grads: bool = True,
                 dx=None, lddx: int = 0, stride_dx: int = 0, dx_col0: int = 0, dx_cols: int = 0, member0: int = 0):
        """d_head: fp32 [E, rows, head_out] gradient w.r.t. the (pre-activation) head output.
        Accumulates dW/db into arena.grads (RED) when weight_grads; optionally writes the gradient
        w.r.t. input columns [dx_col0, dx_col0+dx_cols) into dx (fp32 [E, rows, dx_cols])."""
        L = lib()
        E = ctx.members
        ms = self.arena.member_size
        n, feat = self.head_out, self.feat
        ldh, sdh = n, rows * n
        nl = len(self.hidden)
        if self.precision == "fp32":
            last = ctx.acts[-1]
            dcur = ctx.dz[nl - 1]
            ar = ctx.rows   # member stride of the saved activations (rows <= ar: backward over a prefix of the rows)
            # weight gradients on a side stream: dW_l only needs dZ_l, the chain dZ_l -> dZ_{l-1} does not need dW_l
            side = self._wgrad_stream() if weight_grads and self.wgrad_side else None
            ws = side if side is not None else stream
            if self.wide_head:
                _wide_head_backward(self, ctx, rows, E, d_head, _p(last), feat, ar * feat, _p(dcur), feat, rows * feat,
                                    stream, weight_grads, member0)
            else:
                if weight_grads:
                    if side is not None:
                        L.stream_fork(stream, side)
                    L.head_backward_weight(_p(d_head), ldh, sdh, _p(last), feat, ar * feat,
                                           self._hw("grads", member0), feat, ms, self._hb("grads", member0), ms, rows,
                                           n, feat, E, ws)
                L.head_backward_data(_p(d_head), ldh, sdh, self._hw("params", member0), feat, ms, _p(dcur), feat,
                                     rows * feat, _p(last), feat, ar * feat, rows, n, feat, E, stream)
            for i in range(nl - 1, -1, -1):
                h = self.hidden[i]
                d_in = self.hidden[i - 1] if i > 0 else self.in_dim
                if i > 0:
                    inp, ldi, si = _p(ctx.acts[i - 1]), d_in, ar * d_in
                else:
                    inp, ldi, si = _p(x), ldx, 0
                if weight_grads:
                    if side is not None:
                        L.stream_fork(stream, side)   # dZ_i is complete on the main stream
                    L.linear_backward_weight(_p(dcur), h, rows * h, inp, ldi, si, self._w("grads", i, member0), d_in,
                                             ms, self._b("grads", i, member0), ms, rows, h, d_in, E, ws)
                if i > 0:
                    dnext = ctx.dz[i - 1]
                    L.linear_backward_data(_p(dcur), h, rows * h, self._w("params", i, member0), d_in, ms, _p(dnext),
                                           d_in, rows * d_in, inp, ldi, si, rows, h, d_in, E, stream)
                    dcur = dnext
                elif dx is not None:
                    L.linear_backward_data(_p(dcur), h, rows * h, self._w("params", 0, member0) + 4 * dx_col0, d_in,
                                           ms, _p(dx), lddx, stride_dx, None, 0, 0, rows, h, dx_cols, E, stream)
            if side is not None:
                L.stream_join(stream, side)
            return
        # ---- bf16 mode.  dgrad: C = A B^T over K-major operands (dZ_l, W_l^T shadow); wgrad: C += A^T B with
        # both operands row-major (MN-major UMMA tiles): dW_l = dZ_l^T H_{l-1}, reduction over the minibatch rows
        sms = self.shadow_member
        if self.fused_ok and n <= 16:
            # ONE persistent launch for the whole data-gradient chain (+ bias / head gradients), then one
            # MN-major weight-gradient GEMM per layer over the stored dZ_l and saved H_{l-1}
            import ctypes

            dims = (ctypes.c_int * (nl + 1))(self.in_dim, *self.hidden)
            wp = (ctypes.c_void_p * nl)(*[self._sw("params", i, member0)[0] for i in range(nl)])
            ldw = (ctypes.c_int64 * nl)(*[self._sw("params", i, member0)[1] for i in range(nl)])
            ap = (ctypes.c_void_p * nl)(*[_p(t) for t in ctx.hb])
            lda = (ctypes.c_int64 * nl)(*[t.shape[2] for t in ctx.hb])
            sa = (ctypes.c_int64 * nl)(*[t.shape[1] * t.shape[2] for t in ctx.hb])
            if weight_grads:
                dzp = (ctypes.c_void_p * nl)(*[_p(t) for t in ctx.dz])
                lddz = (ctypes.c_int64 * nl)(*[t.shape[2] for t in ctx.dz])
                sdz = (ctypes.c_int64 * nl)(*[t.shape[1] * t.shape[2] for t in ctx.dz])
                dbp = (ctypes.c_void_p * nl)(*[self._b("grads", i, member0) for i in range(nl)])
            else:
                dzp = lddz = sdz = dbp = None
            dh16 = None
            if weight_grads and n > 1:
                if getattr(ctx, "dh16", None) is None:
                    ctx.dh16 = torch.zeros(E, ctx.rows, 16, dtype=torch.bfloat16, device=self.device)
                dh16 = ctx.dh16
            L.mlp_backward_bf16(rows, E, nl, dims, wp, ldw, sms, ap, lda, sa, dzp, lddz, sdz, _p(d_head),
                                self._hw("params", member0), ms, n, dbp,
                                self._hw("grads", member0) if weight_grads else None,
                                self._hb("grads", member0) if weight_grads else None, ms, _p(dh16),
                                _p(dx) if dx is not None else None, lddx, stride_dx, dx_col0, dx_cols, stream)
            if weight_grads:
                # every dW_l = dZ_l^T H_{l-1} of the network in one launch (MN-major operands, split-K RED)
                a_p, a_ld, a_s, b_p, b_ld, b_s, ms_, ns_, o_p, o_ld = [], [], [], [], [], [], [], [], [], []
                for i in range(nl):
                    h, d_in = self.hidden[i], (self.hidden[i - 1] if i > 0 else self.in_dim)
                    dz = ctx.dz[i]
                    a_p.append(_p(dz)); a_ld.append(dz.shape[2]); a_s.append(dz.shape[1] * dz.shape[2])
                    if i > 0:
                        prev = ctx.hb[i - 1]
                        b_p.append(_p(prev)); b_ld.append(prev.shape[2]); b_s.append(prev.shape[1] * prev.shape[2])
                    else:
                        b_p.append(ctx.x_src[0]); b_ld.append(ctx.x_src[1]); b_s.append(0)
                    ms_.append(h); ns_.append(d_in)
                    o_p.append(self._w("grads", i, member0)); o_ld.append(d_in)
                if dh16 is not None:  # head dW = d_head^T H_{L-1} as one more problem of the same launch
                    last = ctx.hb[-1]
                    a_p.append(_p(dh16)); a_ld.append(16); a_s.append(rows * 16)
                    b_p.append(_p(last)); b_ld.append(last.shape[2]); b_s.append(last.shape[1] * last.shape[2])
                    ms_.append(n); ns_.append(feat)
                    o_p.append(self._hw("grads", member0)); o_ld.append(feat)
                arr = lambda T, v: (T * len(v))(*v)
                L.umma_gemm_tn_batched(len(a_p), arr(ctypes.c_void_p, a_p), arr(ctypes.c_int64, a_ld), arr(ctypes.c_int64, a_s),
                                       arr(ctypes.c_void_p, b_p), arr(ctypes.c_int64, b_ld), arr(ctypes.c_int64, b_s),
                                       arr(ctypes.c_int, ms_), arr(ctypes.c_int, ns_), rows, E,
                                       arr(ctypes.c_void_p, o_p), arr(ctypes.c_int64, o_ld), ms, stream)
            return
        last = ctx.hb[-1]
        lf = _a8(feat)
        dcur = ctx.dz[nl - 1]
        # per-layer launches (layers wider than the fused kernels' 256 columns): weight / bias gradients on a side stream
        side = self._wgrad_stream() if weight_grads and self.wgrad_side and not self.wide_head else None
        ws = side if side is not None else stream
        if side is not None:
            L.stream_fork(stream, side)
        if self.wide_head:
            _wide_head_backward(self, ctx, rows, E, d_head, _p(last), lf, rows * lf, _p(dcur), lf, rows * lf, stream,
                                weight_grads, member0)
        else:
            if weight_grads:
                L.head_backward_weight_bf16(_p(d_head), ldh, sdh, _p(last), lf, rows * lf, self._hw("grads", member0),
                                            feat, ms, self._hb("grads", member0), ms, rows, n, feat, E, ws)
            L.head_backward_data_bf16(_p(d_head), ldh, sdh, self._hw("params", member0), feat, ms, _p(dcur), lf,
                                      rows * lf, None, 0, 0, _p(last), lf, rows * lf, rows, n, feat, E, stream)
        for i in range(nl - 1, -1, -1):
            h = self.hidden[i]
            lh = _a8(h)
            d_in = self.hidden[i - 1] if i > 0 else self.in_dim
            ldi = _a8(d_in)
            if weight_grads:
                if side is not None:
                    L.stream_fork(stream, side)   # dZ_i is complete on the main stream
                L.colsum_bf16(_p(dcur), lh, rows * lh, self._b("grads", i, member0), ms, rows, h, E, ws)
                if i > 0:
                    bsrc, ldb, sb = _p(ctx.hb[i - 1]), ldi, rows * ldi
                else:
                    bsrc, ldb, sb = ctx.x_src[0], ctx.x_src[1], 0
                tiles = -(-h // 128) * -(-d_in // 256) * E
                splits = max(1, min(-(-rows // 64), -(-148 // tiles)))
                L.umma_gemm_tn(_p(dcur), lh, rows * lh, bsrc, ldb, sb, h, d_in, rows, E, splits,
                               self._w("grads", i, member0), d_in, ms, 1, ws)
            if i > 0:
                dnext = ctx.dz[i - 1]
                wt, ldwt = self._swt(i, member0)
                # dZ_{i-1} = (dZ_i W_i) * [H_{i-1} > 0]   (B operand = W_i^T [d_in][h])
                L.umma_gemm(_p(dcur), lh, rows * lh, wt, ldwt, sms, rows, d_in, h, E, 1, None, 0, 0, _p(ctx.hb[i - 1]),
                            ldi, rows * ldi, _p(dnext), ldi, rows * ldi, None, 0, 0, None, 0, 0, 0, stream)
                dcur = dnext
            elif dx is not None:
                wt, ldwt = self._swt(0, member0)
                L.umma_gemm(_p(dcur), lh, rows * lh, wt + 2 * dx_col0 * ldwt, ldwt, sms, rows, dx_cols, h, E, 1, None, 0,
                            0, None, 0, 0, None, 0, 0, None, 0, 0, _p(dx), lddx, stride_dx, 0, stream)
        if side is not None:
            L.stream_join(stream, side)

    # ------------------------------------------------------------------ optimizer
    def adam(self, lr: float, stream: int, betas=(0.9, 0.999), eps=1e-8, tau: Optional[float] = None, peer=None):
        """peer = (PeerExchange, grads pointer table, flag index, block-counter pointer, epoch pointer): the gradient
        all-reduce over NVLink peer memory is fused into the optimizer pass (no NCCL call)."""
        a = self.arena
        sync = tau is not None and a.target is not None
        if peer is not None:
            import ctypes

            px, gptrs, fidx, cptr, eptr = peer[:5]
            small = peer[5] if len(peer) > 5 else None   # (tensor, channel): sums riding along with the exchange
            gred, cptr2 = (peer[6], peer[7]) if len(peer) > 7 else (None, None)   # two-shot exchange buffers
            shadow = self.precision == "bf16" and self.fused_ok and self.head_out <= 16
            seg, nseg = None, 0
            if shadow:
                t = self._table
                nseg = t.shape[0]
                seg = (ctypes.c_int64 * (5 * nseg))(*[int(v) for r in t.tolist() for v in r[:5]])
            lib().adam_step_peer(_p(a.params), _p(a.exp_avg), _p(a.exp_avg_sq), _p(a.target) if sync else None, a.size,
                                 _p(a.step), lr, betas[0], betas[1], eps, tau if sync else 0.0,
                                 _p(self.shadow) if shadow else None,
                                 _p(self.shadow_target) if shadow and sync and self.shadow_target is not None else None,
                                 seg, nseg, a.member_size, self.shadow_member if shadow else 0, gptrs, px.flags_ptrs,
                                 px.world, px.rank, fidx, eptr, cptr, _p(small[0]) if small else None,
                                 small[0].numel() if small else 0, px.xchg_ptrs if small else None,
                                 small[1] if small else 0, gred, cptr2, stream)
            if not shadow:
                self.refresh_shadow("params", stream)
                if sync:
                    self.refresh_shadow("target", stream)
            return
        if self.precision == "bf16" and self.fused_ok and self.head_out <= 16:
            # the fused kernels only read the row-major W shadows: refresh them inside the Adam pass
            import ctypes

            t = self._table
            seg = (ctypes.c_int64 * (5 * t.shape[0]))(*[int(v) for r in t.tolist() for v in (r[0], r[1], r[2], r[3], r[4])])
            lib().adam_step_shadow(_p(a.params), _p(a.grads), _p(a.exp_avg), _p(a.exp_avg_sq),
                                   _p(a.target) if sync else None, a.size, _p(a.step), lr, betas[0], betas[1], eps,
                                   tau if sync else 0.0, _p(self.shadow),
                                   _p(self.shadow_target) if sync and self.shadow_target is not None else None, seg,
                                   t.shape[0], a.member_size, self.shadow_member, stream)
            return
        lib().adam_step(_p(a.params), _p(a.grads), _p(a.exp_avg), _p(a.exp_avg_sq),
                        _p(a.target) if sync else None, a.size, _p(a.step), lr, betas[0], betas[1], eps,
                        tau if sync else 0.0, 1, stream)
        self.refresh_shadow("params", stream)
        if sync:
            self.refresh_shadow("target", stream)


NATURE_FILTERS = [(32, 8, 4), (64, 4, 2), (64, 3, 1)]  # PixelEncoderFactory defaults (models/encoders.py:100-120)


class ConvCtx:
    """Workspace of one ConvNet forward(/backward): patch matrices and NHWC activations per layer (fp32, or bf16
    with 16-byte padded rows in tensor-core mode)."""

    def __init__(self, net: "ConvNet", images: int, members: int, train: bool):
        self.images, self.members, self.train = images, members, train
        dev, E = net.device, members
        bf = net.precision == "bf16"
        dt = torch.bfloat16 if bf else torch.float32
        pad = _a8 if bf else (lambda n: n)
        self.patches, self.acts = [], []
        for i, (C, H, W, oc, k, s, OH, OW) in enumerate(net.layers):
            rows, K = images * OH * OW, C * k * k
            self.patches.append(torch.zeros(1 if i == 0 else E, rows, pad(K), dtype=dt, device=dev))
            self.acts.append(torch.zeros(E, rows, pad(oc), dtype=dt, device=dev))
        if train:
            self.dacts = [torch.zeros(E, images * OH * OW, pad(oc), dtype=dt, device=dev)
                          for (_, _, _, oc, _, _, OH, OW) in net.layers]
            kmax = max(images * OH * OW * C * k * k for (C, _, _, _, k, _, OH, OW) in net.layers[1:])
            self.dpatch = torch.zeros(E * kmax, dtype=torch.float32, device=dev)


class ConvNet:
    """Nature-DQN pixel encoder + narrow head for all ensemble members (PixelEncoder + DiscreteMeanQFunction,
    d3rlpy/models/torch/encoders.py:43-162, q_functions/mean_q_function.py:13-24).  Every layer — the three
    convolutions and the flatten+fc — is `im2col` followed by the batched dense-layer GEMM; activations
    are NHWC.  Input is the uint8 NCHW frame stack straight from the gather kernel."""

    def __init__(self, obs_shape, heads: Sequence[Tuple[str, int]], members: int, device, feature_size: int = 512,
                 filters=None, member_key: str = "{name}", with_target: bool = False,
                 seed_gen: Optional[torch.Generator] = None, precision: str = "fp32", input_divisor: float = 1.0):
        assert precision in ("fp32", "bf16")
        C, H, W = obs_shape
        self.obs_shape, self.members, self.device = (C, H, W), members, device
        self.precision, self.input_divisor = precision, input_divisor
        self.heads = list(heads)
        self.head_out = sum(n for _, n in heads)
        filters = list(filters) if filters is not None else NATURE_FILTERS
        self.layers = []
        entries = []
        for l, (oc, k, s) in enumerate(filters):
            OH, OW = (H - k) // s + 1, (W - k) // s + 1
            self.layers.append((C, H, W, oc, k, s, OH, OW))
            entries.append((f"_encoder._convs.{l}.weight", (oc, C, k, k)))
            entries.append((f"_encoder._convs.{l}.bias", (oc,)))
            C, H, W = oc, OH, OW
        assert H == W, "square feature maps only"
        self.layers.append((C, H, W, feature_size, H, 1, 1, 1))  # flatten + fc as a whole-map convolution
        entries.append(("_encoder._fc.weight", (feature_size, C * H * W)))
        entries.append(("_encoder._fc.bias", (feature_size,)))
        self._names = [f"_encoder._convs.{l}" for l in range(len(filters))] + ["_encoder._fc"]
        self.feat = feature_size
        entries.append(("__head.weight", (self.head_out, feature_size)))
        entries.append(("__head.bias", (self.head_out,)))
        exports, r0 = [], 0
        for name, n in self.heads:
            exports.append((f"{name}.weight", "__head.weight", r0, n))
            exports.append((f"{name}.bias", "__head.bias", r0, n))
            r0 += n
        self.arena = ParamArena(entries, members, device, with_target=with_target, member_key=member_key,
                                exports=exports)
        self.wide_head = self.head_out > HEAD_NARROW_MAX
        self._ctx: Dict[str, ConvCtx] = {}
        self._init_params(seed_gen)
        if precision == "bf16":
            self._build_shadow(with_target)

    # ---- bf16 shadows: W [oc][ld8(K)] (forward B operand) and W^T [K][ld8(oc)] (dgrad B operand), K-major both
    def _build_shadow(self, with_target: bool):
        rows, off = [], 0
        self._sh_w, self._sh_wt = [], []
        for i, (C, H, W, oc, k, s, OH, OW) in enumerate(self.layers):
            K = C * k * k
            ldk, ldn = _a8(K), _a8(oc)
            w_off = off
            off += oc * ldk
            wt_off = off
            off += K * ldn
            self._sh_w.append((w_off, ldk))
            self._sh_wt.append((wt_off, ldn))
            rows.append([self.arena.offsets[self._names[i] + ".weight"], oc, K, w_off, ldk, wt_off, ldn])
        if self.wide_head:
            d, n = self.feat, self.head_out
            ldk, ldn = _a8(d), _a8(n)
            self._head_shadow = len(self._sh_w)
            self._sh_w.append((off, ldk))
            self._sh_wt.append((off + n * ldk, ldn))
            rows.append([self.arena.offsets["__head.weight"], n, d, off, ldk, off + n * ldk, ldn])
            off += n * ldk + d * ldn
        self.shadow_member = _a8(off)
        self._table = torch.tensor(rows, dtype=torch.int64)
        n = self.shadow_member * self.members
        self.shadow = torch.zeros(n, dtype=torch.bfloat16, device=self.device)
        self.shadow_target = torch.zeros(n, dtype=torch.bfloat16, device=self.device) if with_target else None

    def _sw(self, which, i):
        base = self.shadow if which == "params" else self.shadow_target
        return base.data_ptr() + 2 * self._sh_w[i][0], self._sh_w[i][1]

    def _swt(self, i):
        return self.shadow.data_ptr() + 2 * self._sh_wt[i][0], self._sh_wt[i][1]

    def _init_params(self, gen):
        a = self.arena
        with torch.no_grad():
            for e in range(self.members):
                for name, shape in a.shapes.items():
                    wshape = shape if len(shape) > 1 else a.shapes[name[:-4] + "weight"]
                    fan_in = 1
                    for s in wshape[1:]:
                        fan_in *= s
                    bound = 1.0 / math.sqrt(fan_in)
                    a.view(name, e).copy_((torch.rand(shape, generator=gen) * 2 - 1) * bound)
        a.sync_target_from_params()

    def refresh_shadow(self, which: str, stream: int):
        if self.precision != "bf16":
            return
        src = self.arena.params if which == "params" else self.arena.target
        dst = self.shadow if which == "params" else self.shadow_target
        lib().shadow_weights(_p(src), self.arena.member_size, _p(dst), self.shadow_member, self._table.data_ptr(),
                             self._table.shape[0], self.members, stream)

    def ctx(self, tag: str, images: int, members: Optional[int] = None, train: bool = True) -> ConvCtx:
        E = members or self.members
        c = self._ctx.get(tag)
        if c is None or c.images != images or c.members != E or c.train != train:
            c = ConvCtx(self, images, E, train)
            self._ctx[tag] = c
        return c

    def _w(self, which, l, member=0):
        return self.arena.addr(which, self._names[l] + ".weight", member)

    def _b(self, which, l, member=0):
        return self.arena.addr(which, self._names[l] + ".bias", member)

    def forward(self, which: str, x_u8, images: int, ctx: ConvCtx, head_out, stream: int, x_is_u8: bool = True):
        """x_u8: NCHW frame stack [images, C, H, W] (uint8, or fp32 when x_is_u8 is False)."""
        L, E, ms = lib(), ctx.members, self.arena.member_size
        bf = self.precision == "bf16"
        for i, (C, H, W, oc, k, s, OH, OW) in enumerate(self.layers):
            rows, K = images * OH * OW, C * k * k
            p, y = ctx.patches[i], ctx.acts[i]
            ldp, ldy = p.shape[2], y.shape[2]
            if i == 0:
                L.im2col(_p(x_u8), 1 if x_is_u8 else 0, 0, C * H * W, H * W, W, 1, _p(p), 1 if bf else 0, ldp, 0, images, C,
                         H, W, k, s, self.input_divisor, 1, stream)
                sx = 0
            else:
                prev = ctx.acts[i - 1]
                ldc = prev.shape[2]  # NHWC rows of ldc (>= C) elements
                L.im2col(_p(prev), 2 if bf else 0, prev.shape[1] * ldc, H * W * ldc, 1, W * ldc, ldc, _p(p), 1 if bf else 0,
                         ldp, p.shape[1] * ldp, images, C, H, W, k, s, 1.0, E, stream)
                sx = p.shape[1] * ldp
            if bf:
                wptr, ldw = self._sw(which, i)
                L.umma_gemm(_p(p), ldp, sx, wptr, ldw, self.shadow_member, rows, oc, K, E, 1, self._b(which, i), ms, 1,
                            None, 0, 0, _p(y), ldy, y.shape[1] * ldy, None, 0, 0, None, 0, 0, 0, stream)
            else:
                L.linear_forward(_p(p), K, sx, self._w(which, i), K, ms, self._b(which, i), ms, _p(y), oc,
                                 y.shape[1] * oc, rows, oc, K, E, 1, stream)
        if head_out is not None and self.wide_head:
            last = ctx.acts[-1]
            _wide_head_forward(self, which, _p(last), last.shape[2], last.shape[1] * last.shape[2], images, E, head_out,
                               stream)
        elif head_out is not None:
            last, n, d = ctx.acts[-1], self.head_out, self.feat
            hf = L.head_forward_bf16 if bf else L.head_forward
            hf(_p(last), last.shape[2], last.shape[1] * last.shape[2], self.arena.addr(which, "__head.weight"), d, ms,
               self.arena.addr(which, "__head.bias"), ms, _p(head_out), n, images * n, images, n, d, E, 0, stream)

    def backward(self, images: int, ctx: ConvCtx, d_head, stream: int):
        """d_head: fp32 [E, images, head_out].  Accumulates every dW/db into arena.grads."""
        L, E, ms = lib(), ctx.members, self.arena.member_size
        n, d = self.head_out, self.feat
        bf = self.precision == "bf16"
        last = ctx.acts[-1]
        ldl = last.shape[2]
        dcur = ctx.dacts[-1]
        ldd = dcur.shape[2]
        hw_g, hb_g = self.arena.addr("grads", "__head.weight"), self.arena.addr("grads", "__head.bias")
        # weight / bias gradients on a side stream beside the data-gradient chain (dW_l needs dY_l only; every layer
        # has its own dY buffer)
        # (bf16 mode only: measured on c4, the fp32-mode weight gradients — 3xTF32 launches split to fill the machine —
        # slow the data-gradient chain down more than the overlap gains: 507 -> 655 us)
        side = self._wgrad_stream() if self.wgrad_side and bf and not self.wide_head else None
        ws = side if side is not None else stream
        if side is not None:
            L.stream_fork(stream, side)
        if self.wide_head:
            _wide_head_backward(self, ctx, images, E, d_head, _p(last), ldl, last.shape[1] * ldl, _p(dcur), ldd,
                                dcur.shape[1] * ldd, stream)
        elif bf:
            L.head_backward_weight_bf16(_p(d_head), n, images * n, _p(last), ldl, last.shape[1] * ldl, hw_g, d, ms, hb_g,
                                        ms, images, n, d, E, ws)
            L.head_backward_data_bf16(_p(d_head), n, images * n, self.arena.addr("params", "__head.weight"), d, ms,
                                      _p(dcur), ldd, dcur.shape[1] * ldd, None, 0, 0, _p(last), ldl, last.shape[1] * ldl,
                                      images, n, d, E, stream)
        else:
            L.head_backward_weight(_p(d_head), n, images * n, _p(last), d, last.shape[1] * d, hw_g, d, ms, hb_g, ms,
                                   images, n, d, E, ws)
            L.head_backward_data(_p(d_head), n, images * n, self.arena.addr("params", "__head.weight"), d, ms, _p(dcur),
                                 d, dcur.shape[1] * d, _p(last), d, last.shape[1] * d, images, n, d, E, stream)
        for i in range(len(self.layers) - 1, -1, -1):
            C, H, W, oc, k, s, OH, OW = self.layers[i]
            rows, K = images * OH * OW, C * k * k
            p = ctx.patches[i]
            ldp, ldd = p.shape[2], dcur.shape[2]
            sx = 0 if i == 0 else p.shape[1] * ldp
            if side is not None:
                L.stream_fork(stream, side)   # dY_i is complete on the main stream
            if bf:
                L.colsum_bf16(_p(dcur), ldd, dcur.shape[1] * ldd, self._b("grads", i), ms, rows, oc, E, ws)
                tiles = -(-oc // 128) * -(-K // 128) * E
                splits = max(1, min(-(-rows // 64), -(-296 // tiles)))
                # dW[oc][K] += dY^T patches: both operands row-major => MN-major tensor-core tiles
                L.umma_gemm_tn(_p(dcur), ldd, dcur.shape[1] * ldd, _p(p), ldp, sx, oc, K, rows, E, splits,
                               self._w("grads", i), K, ms, 1, ws)
            else:
                L.linear_backward_weight(_p(dcur), oc, dcur.shape[1] * oc, _p(p), K, sx, self._w("grads", i), K, ms,
                                         self._b("grads", i), ms, rows, oc, K, E, ws)
            if i == 0:
                break
            sdp = p.shape[1] * K
            if bf:
                wt, ldwt = self._swt(i)
                # dpatch[rows][K] = dY[rows][oc] . W[oc][K]   (B operand = W^T [K][oc], K-major)
                L.umma_gemm(_p(dcur), ldd, dcur.shape[1] * ldd, wt, ldwt, self.shadow_member, rows, K, oc, E, 1, None, 0,
                            0, None, 0, 0, None, 0, 0, None, 0, 0, _p(ctx.dpatch), K, sdp, 0, stream)
            else:
                L.linear_backward_data(_p(dcur), oc, dcur.shape[1] * oc, self._w("params", i), K, ms, _p(ctx.dpatch), K,
                                       sdp, None, 0, 0, rows, oc, K, E, stream)
            prev, dprev = ctx.acts[i - 1], ctx.dacts[i - 1]
            L.col2im(_p(ctx.dpatch), K, sdp, _p(prev), 1 if bf else 0, prev.shape[2], prev.shape[1] * prev.shape[2],
                     _p(dprev), dprev.shape[2], dprev.shape[1] * dprev.shape[2], images, C, H, W, k, s, E, stream)
            dcur = dprev
        if side is not None:
            L.stream_join(stream, side)

    wgrad_side = True

    def _wgrad_stream(self) -> int:
        if getattr(self, "_wg_side", None) is None:
            self._wg_side = torch.cuda.Stream(device=self.device)
        return self._wg_side.cuda_stream

    def adam(self, lr: float, stream: int, betas=(0.9, 0.999), eps=1e-8, tau: Optional[float] = None):
        a = self.arena
        lib().adam_step(_p(a.params), _p(a.grads), _p(a.exp_avg), _p(a.exp_avg_sq),
                        _p(a.target) if tau is not None else None, a.size, _p(a.step), lr, betas[0], betas[1], eps,
                        tau if tau is not None else 0.0, 1, stream)
        self.refresh_shadow("params", stream)
        if tau is not None and a.target is not None:
            self.refresh_shadow("target", stream)
