"""Phase stamps (clock64) of the fused MLP forward kernel for the c2 shapes: where the per-launch latency goes.
stamp 0 = after setup (barriers, TMEM alloc), 1 = after griddepcontrol.wait, 2+2g = accumulator of layer g ready,
3+2g = epilogue of layer g done, 14 = last store drained, 15 = teardown."""
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3rlpy_b200._lib import lib  # noqa: E402

L, dev = lib(), torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
a8 = lambda v: (v + 7) // 8 * 8
for rows, E, in_dim, hidden, n_head, save in [(256, 2, 23, [256, 256, 256], 1, True), (256, 2, 23, [256, 256, 256], 1, False), (512, 1, 17, [256, 256, 256], 12, True),
                                              (7936, 2, 23, [256, 256, 256], 1, True), (15872, 2, 23, [256, 256, 256], 1, True)]:
    g = torch.Generator().manual_seed(0)
    dims = [in_dim] + hidden
    x = torch.randn(rows, a8(in_dim), generator=g).to(torch.bfloat16).to(dev)
    w_off, off = [], 0
    for k, n in zip(dims[:-1], dims[1:]):
        w_off.append(off)
        off += n * a8(k)
    sms = a8(off)
    shadow = (torch.randn(E, sms, generator=g) / 16).to(torch.bfloat16).to(dev)
    feat = hidden[-1]
    nb = sum(hidden)
    ms = nb + n_head * feat + n_head
    arena = (torch.randn(E, ms, generator=g) * 0.05).to(dev)
    b_off = [sum(hidden[:i]) for i in range(len(hidden))]
    acts = [torch.zeros(E, rows, a8(n), dtype=torch.bfloat16, device=dev) for n in hidden]
    out = torch.zeros(E, rows, n_head, device=dev)
    units = -(-rows // 128) * E
    grid = min(units, 148)
    dbg = torch.zeros(grid * 16, dtype=torch.int64, device=dev)
    arr = lambda T, v: (T * len(v))(*v)

    def run():
        L.mlp_forward_bf16(x.data_ptr(), a8(in_dim), 0, rows, E, len(hidden), arr(ctypes.c_int, dims),
                           arr(ctypes.c_void_p, [shadow.data_ptr() + 2 * o for o in w_off]),
                           arr(ctypes.c_int64, [a8(k) for k in dims[:-1]]), sms,
                           arr(ctypes.c_void_p, [arena.data_ptr() + 4 * o for o in b_off]), ms,
                           arr(ctypes.c_void_p, [a.data_ptr() for a in acts]) if save else None,
                           arr(ctypes.c_int64, [a.shape[2] for a in acts]) if save else None,
                           arr(ctypes.c_int64, [a.shape[1] * a.shape[2] for a in acts]) if save else None,
                           arena.data_ptr() + 4 * nb, arena.data_ptr() + 4 * (nb + n_head * feat), ms, n_head, 0,
                           out.data_ptr(), 0, st)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    L.mlp_set_debug(dbg.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    L.mlp_set_debug(None)
    d = dbg.view(grid, 16).cpu()
    cta = d[0]
    base = int(cta[0])
    rel = [(int(v) - base) if int(v) else None for v in cta]
    print(f"rows={rows} E={E} heads={n_head} units={units} grid={grid} event_us={1e3 * e0.elapsed_time(e1):.1f}")
    print("  CTA0 cycles since setup:", rel)
    print("  teardown stamp min/max over CTAs (cycles since their own setup):",
          int((d[:, 15] - d[:, 0]).min()), int((d[:, 15] - d[:, 0]).max()))


# ---- backward kernel: stamp 0 setup, 1 after griddepcontrol.wait, 2 head constants staged, then per produced dZ_l
# (top layer first) 3+3s inputs ready (accumulator + mask tile), 4+3s dZ_l written, 5+3s column passes done; 15 end
print("== mlp_backward_bf16")
for rows, E, in_dim, hidden, n_head, wg, dxr in [(7936, 2, 23, [256, 256, 256], 1, True, None),
                                                 (256, 2, 23, [256, 256, 256], 1, False, (17, 6)),
                                                 (256, 1, 17, [256, 256, 256], 12, True, None)]:
    g = torch.Generator().manual_seed(1)
    dims = [in_dim] + hidden
    nl = len(hidden)
    w_off, off = [], 0
    for k, n in zip(dims[:-1], dims[1:]):
        w_off.append(off)
        off += n * a8(k)
    sms = a8(off)
    shadow = (torch.randn(E, sms, generator=g) / 16).to(torch.bfloat16).to(dev)
    feat = hidden[-1]
    acts = [torch.relu(torch.randn(E, rows, a8(n), generator=g)).to(torch.bfloat16).to(dev) for n in hidden]
    b_off = [sum(hidden[:i]) for i in range(nl)]
    hw_off, hb_off = sum(hidden), sum(hidden) + n_head * feat
    ms = (hb_off + n_head + 3) // 4 * 4
    params = (torch.randn(E, ms, generator=g) * 0.05).to(dev)
    grads = torch.zeros(E, ms, device=dev)
    d_head = (torch.randn(E, rows, n_head, generator=g) / rows).to(dev)
    dz = [torch.zeros(E, rows, a8(n), dtype=torch.bfloat16, device=dev) for n in hidden]
    dh16 = torch.zeros(E, rows, 16, dtype=torch.bfloat16, device=dev)
    dx = torch.zeros(E, rows, dxr[1], device=dev) if dxr else None
    units = -(-rows // 128) * E
    grid = min(units, 148)
    dbg = torch.zeros(grid * 16, dtype=torch.int64, device=dev)
    arr = lambda T, v: (T * len(v))(*v)

    def run_b():
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
                            dh16.data_ptr() if (wg and n_head > 1) else None, dx.data_ptr() if dxr else None,
                            dxr[1] if dxr else 0, rows * dxr[1] if dxr else 0, dxr[0] if dxr else 0, dxr[1] if dxr else 0, st)

    for _ in range(3):
        run_b()
    torch.cuda.synchronize()
    L.mlp_set_debug(dbg.data_ptr())
    run_b()
    torch.cuda.synchronize()
    L.mlp_set_debug(None)
    d = dbg.view(grid, 16).cpu()
    base = int(d[0][0])
    print(f"rows={rows} E={E} heads={n_head} wgrad={wg} dx={dxr} units={units}")
    print("  CTA0 cycles since setup:", [(int(v) - base) if int(v) else None for v in d[0]])
    print("  end stamp min/max over CTAs:", int((d[:, 15] - d[:, 0]).min()), int((d[:, 15] - d[:, 0]).max()))
