"""Phase timing of the tcgen05 GEMM kernel (clock64 stamps per CTA) + CUDA-event durations for a sweep
of shapes.  Run on the GPU box:  python profiles/umma_phase_probe.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3rlpy_b200._lib import lib  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
side = torch.cuda.Stream()
st = side.cuda_stream


def run(M, N, K, E, mode, reps=20):
    a = torch.randn(E, M, K, device=dev).to(torch.bfloat16)
    b = (torch.randn(E, N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(E, N, device=dev)
    mask = torch.randn(E, M, N, device=dev).to(torch.bfloat16)
    out = torch.zeros(E, M, N, dtype=torch.bfloat16, device=dev)
    Mp = (M + 7) // 8 * 8
    out_t = torch.zeros(E, N, Mp, dtype=torch.bfloat16, device=dev)
    out_f = torch.zeros(E, M, N, device=dev)
    p = lambda t: None if t is None else t.data_ptr()
    kw = dict(bias=None, relu=0, mask=None, ob=None, ot=None, of=None, atomic=0)
    if mode == "fwd":
        kw.update(bias=bias, relu=1, ob=out)
    elif mode == "fwd_t":
        kw.update(bias=bias, relu=1, ob=out, ot=out_t)
    elif mode == "dgrad":
        kw.update(mask=mask, ob=out, ot=out_t)
    elif mode == "f32":
        kw.update(of=out_f)
    elif mode == "red":
        kw.update(of=out_f, atomic=1)

    def call():
        L.umma_gemm(p(a), K, M * K, p(b), K, N * K, M, N, K, E, 1, p(kw["bias"]), N, kw["relu"], p(kw["mask"]), N, M * N,
                    p(kw["ob"]), N, M * N, p(kw["ot"]), Mp, N * Mp, p(kw["of"]), N, M * N, kw["atomic"], st)

    n_cta = 4 * (-(-M // 128)) * E * 2
    dbg = torch.zeros(n_cta * 8, dtype=torch.int64, device=dev)
    import ctypes

    torch.cuda.synchronize()
    call()
    torch.cuda.synchronize()
    # GPU-side time: capture `reps` back-to-back launches in a CUDA graph and time its replay
    L.graph_begin(st)
    for _ in range(reps):
        call()
    g, nn = ctypes.c_void_p(), ctypes.c_int()
    L.graph_end(st, ctypes.byref(g), ctypes.byref(nn))
    L.graph_launch(g.value, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    L.graph_launch(g.value, st)
    e1.record(side)
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    L.graph_destroy(g.value)
    L.umma_set_debug(dbg.data_ptr())
    call()
    torch.cuda.synchronize()
    L.umma_set_debug(None)
    d = dbg.view(n_cta, 8).cpu()
    d = d[d[:, 0] != 0]
    n_cta = d.shape[0]
    rel = (d[:, 1:] - d[:, :1]).float()
    med = rel.median(0).values.tolist()
    names = ["setup", "tma_issued", "stage0_landed", "mma_issued", "acc_ready", "epi_done", "exit"]
    print(f"{mode:6s} M={M:5d} N={N:4d} K={K:5d} E={E} ctas={n_cta:4d}  {us:7.1f} us/launch   cycles: " +
          "  ".join(f"{n}={int(v)}" for n, v in zip(names, med)))


if __name__ == "__main__":
    for mode in ("f32", "fwd", "fwd_t", "dgrad", "red"):
        run(512, 256, 256, 1, mode)
    for N in (32, 64, 128, 256):
        run(512, N, 256, 1, "fwd")
    for K in (64, 256, 1024, 4096):
        run(512, 256, K, 1, "fwd")
    run(7936, 256, 256, 2, "fwd_t")
    run(7936, 256, 256, 2, "dgrad")
