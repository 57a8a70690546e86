"""Graph-timed latency of small bf16 umma_gemm launches with the cluster split-K configuration on / off."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200._lib import lib  # noqa: E402
import profiles.r2.tc32_bench_lib as tb  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
for (M, N, K, E) in [(256, 750, 752, 1), (256, 300, 400, 2), (256, 400, 24, 2), (32, 512, 3136, 1), (512, 256, 256, 1),
                     (100, 750, 752, 1), (25600, 750, 752, 1)]:
    a = torch.randn(E, M, K, device=dev).bfloat16()
    b = (torch.randn(E, N, K, device=dev) / math.sqrt(K)).bfloat16()
    bias = torch.randn(E, N, device=dev)
    ldn = (N + 7) // 8 * 8
    out = torch.zeros(E, M, ldn, dtype=torch.bfloat16, device=dev)
    row = f"M={M:6d} N={N:4d} K={K:5d} E={E} "
    for cluster in (1, 0):
        L.umma_set_cluster(cluster)
        t = tb.timeit(lambda st: L.umma_gemm(a.data_ptr(), K, M * K, b.data_ptr(), K, N * K, M, N, K, E, 1, bias.data_ptr(), N,
                                             1, None, 0, 0, out.data_ptr(), ldn, M * ldn, None, 0, 0, None, 0, 0, 0, st))
        row += f"| cluster {cluster}: {t:6.1f} us "
    L.umma_set_cluster(1)
    print(row, flush=True)
