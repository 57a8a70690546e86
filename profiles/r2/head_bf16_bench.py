"""Graph-timed bf16 head forward at the BCQ shapes: 16-byte-chunk kernel (even weight leading dimension) against the
element-wise kernel (forced by an odd leading dimension)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200._lib import lib  # noqa: E402
import profiles.r2.tc32_bench_lib as tb  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
for (M, K, N, E) in [(25600, 750, 6, 1), (25600, 300, 6, 1), (25600, 300, 1, 2), (256, 750, 6, 1)]:
    ld = (K + 7) // 8 * 8
    x = torch.randn(E, M, ld, device=dev).bfloat16()
    b = torch.randn(E, N, device=dev)
    y = torch.zeros(E, M, N, device=dev)
    row = f"M={M:6d} K={K:4d} N={N:2d} E={E} "
    for ldw in (K, K + 1):
        w = torch.randn(E, N, ldw, device=dev) / math.sqrt(K)
        t = tb.timeit(lambda st: L.head_forward_bf16(x.data_ptr(), ld, M * ld, w.data_ptr(), ldw, N * ldw, b.data_ptr(), N,
                                                     y.data_ptr(), N, M * N, M, N, K, E, 0, st))
        row += f"| ldw {ldw}: {t:6.1f} us "
    print(row, flush=True)
