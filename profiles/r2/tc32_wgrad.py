"""Weight-gradient split heuristics of the 3xTF32 GEMM (variant bits 16-17: 0 = default (two CTAs per SM), 1 = ceil(148 / tiles)
splits, 2 = at most one CTA per SM, 3 = four per SM)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200._lib import lib  # noqa: E402
import profiles.r2.tc32_bench_lib as tb  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
for (M, N, K, E) in [(7936, 256, 256, 2), (7936, 256, 23, 2), (256, 256, 256, 2), (256, 256, 17, 1), (81920, 256, 256, 10), (81920, 256, 119, 10)]:
    x = torch.randn(E, M, K, device=dev)
    y = torch.randn(E, M, N, device=dev)
    dw = torch.zeros(E, N, K, device=dev)
    db = torch.zeros(E, N, device=dev)
    row = f"M={M:6d} N={N} K={K:4d} E={E:2d} "
    for mode in (0, 1, 2, 3):
        L.tc32_set_variant(mode << 16)
        tw = tb.timeit(lambda st: L.linear_backward_weight(y.data_ptr(), N, M * N, x.data_ptr(), K, M * K, dw.data_ptr(), K,
                                                           N * K, db.data_ptr(), N, M, N, K, E, st))
        row += f"| mode {mode}: {tw:7.1f} us "
    L.tc32_set_variant(0)
    print(row, flush=True)
