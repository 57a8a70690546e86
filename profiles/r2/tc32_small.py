"""Latency of the small (256 ... 1024-row) 3xTF32 GEMM launches of a batch-256 update under different cluster
split-K sizes / tile widths (forced through d3b_tc32_set_variant: bits 8-11 = cluster size, bits 12-13 = BN index)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200._lib import lib  # noqa: E402
import profiles.r2.tc32_bench_lib as tb  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
for (M, N, K, E) in [(256, 256, 256, 2), (512, 256, 256, 1), (256, 256, 23, 2), (1024, 256, 256, 1), (2560, 256, 256, 2)]:
    x = torch.randn(E, M, K, device=dev)
    w = torch.randn(E, N, K, device=dev) / K ** 0.5
    b = torch.randn(E, N, device=dev)
    y = torch.empty(E, M, N, device=dev)
    dx = torch.empty(E, M, K, device=dev)
    row = f"M={M:5d} N={N} K={K:4d} E={E} "
    for tag, var in (("auto", 0), ("S1", 1 << 8), ("S2/64", (2 << 8) | (2 << 12)), ("S4/64", (4 << 8) | (2 << 12)),
                     ("S8/64", (8 << 8) | (2 << 12)), ("S4/32", (4 << 8) | (1 << 12)), ("S4/128", (4 << 8) | (3 << 12)),
                     ("S8/128", (8 << 8) | (3 << 12))):
        if K < 64 and var not in (0, 1 << 8):
            continue
        L.tc32_set_variant(var)
        tf = tb.timeit(lambda st: L.linear_forward(x.data_ptr(), K, M * K, w.data_ptr(), K, N * K, b.data_ptr(), N,
                                                   y.data_ptr(), N, M * N, M, N, K, E, 1, st))
        td = tb.timeit(lambda st: L.linear_backward_data(y.data_ptr(), N, M * N, w.data_ptr(), K, N * K, dx.data_ptr(), K,
                                                         M * K, x.data_ptr(), K, M * K, M, N, K, E, st))
        row += f"| {tag} f {tf:5.1f} d {td:5.1f} "
    L.tc32_set_variant(0)
    print(row, flush=True)
