"""Which phase bounds the 3xTF32 mainloop: time of the c5-sized forward GEMM with one phase knocked out."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200._lib import lib  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
M, N, K, E = 81920, 256, 256, 2
x = torch.randn(E, M, K, device=dev)
w = torch.randn(E, N, K, device=dev) / K ** 0.5
y = torch.empty(E, M, N, device=dev)
f = lambda: L.linear_forward(x.data_ptr(), K, M * K, w.data_ptr(), K, N * K, None, 0, y.data_ptr(), N, M * N, M, N, K, E, 1, st)
for v, name in [(0, "full"), (1, "no global loads"), (2, "no smem stores"), (4, "no MMAs"), (3, "no loads, no stores"), (6, "no stores, no MMAs"), (5, "no loads no MMAs"), (7, "nothing")]:
    L.tc32_set_variant(v)
    for _ in range(3):
        f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        f()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / 10 * 1e3
    print(f"{name:22s} {us:8.1f} us  {2.0 * M * N * K * E / us / 1e6:6.1f} TF-equivalent")
L.tc32_set_variant(0)
