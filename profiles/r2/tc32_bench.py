"""Micro-benchmark of the 3xTF32 GEMM at the c2 / c5 layer shapes: CUDA-event time per launch (back-to-back
launches, L2-resident operands), effective fp32-equivalent TFLOP/s, next to the SIMT engine."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200._lib import lib  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=20):
    """n launches captured in ONE CUDA graph (the host cannot enqueue a 5 us kernel every 5 us from Python), replayed
    five times; us per launch."""
    global st
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    old = st
    with torch.cuda.graph(g, stream=side):
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(n):
            fn()
    st = old
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * n) * 1e3  # us


shapes = [(7936, 256, 256, 2), (7936, 256, 23, 2), (7936, 256, 1024, 2), (15872, 256, 256, 2), (512, 256, 256, 1),
          (256, 256, 256, 2), (81920, 256, 256, 10), (1024, 256, 256, 1)]
for (M, N, K, E) in shapes:
    x = torch.randn(E, M, K, device=dev)
    w = torch.randn(E, N, K, device=dev) / K ** 0.5
    b = torch.randn(E, N, device=dev)
    y = torch.empty(E, M, N, device=dev)
    dw = torch.zeros(E, N, K, device=dev)
    db = torch.zeros(E, N, device=dev)
    dx = torch.empty(E, M, K, device=dev)
    fl = 2.0 * M * N * K * E
    row = f"M={M:6d} N={N} K={K:5d} E={E:2d} "
    for eng in (1, 0):
        L.set_fp32_engine(eng)
        tf = timeit(lambda: L.linear_forward(x.data_ptr(), K, M * K, w.data_ptr(), K, N * K, b.data_ptr(), N, y.data_ptr(),
                                             N, M * N, M, N, K, E, 1, st))
        td = timeit(lambda: L.linear_backward_data(y.data_ptr(), N, M * N, w.data_ptr(), K, N * K, dx.data_ptr(), K, M * K,
                                                   x.data_ptr(), K, M * K, M, N, K, E, st))
        tw = timeit(lambda: L.linear_backward_weight(y.data_ptr(), N, M * N, x.data_ptr(), K, M * K, dw.data_ptr(), K,
                                                     N * K, db.data_ptr(), N, M, N, K, E, st))
        row += f"| {'tc32' if eng else 'simt'} fwd {tf:7.1f}us {fl / tf / 1e6:6.1f}TF  dgrad {td:7.1f}us {fl / td / 1e6:6.1f}TF  wgrad {tw:7.1f}us {fl / tw / 1e6:6.1f}TF "
    print(row)
L.set_fp32_engine(1)
