"""clock64 phase stamps of one CTA of the 3xTF32 GEMM (producer thread 0 and the MMA thread, first 6 K blocks)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200._lib import lib  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
for (M, N, K, E) in [(256, 256, 256, 2), (15872, 256, 256, 2)]:
    x = torch.randn(E, M, K, device=dev)
    w = torch.randn(E, N, K, device=dev) / K ** 0.5
    y = torch.empty(E, M, N, device=dev)
    nct = -(-M // 128) * -(-N // 32) * E   # upper bound for any BN
    dbg = torch.zeros(nct, 64, dtype=torch.int64, device=dev)
    f = lambda: L.linear_forward(x.data_ptr(), K, M * K, w.data_ptr(), K, N * K, None, 0, y.data_ptr(), N, M * N, M, N, K, E, 1, st)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    L.tc32_set_debug(dbg.data_ptr())
    f()
    torch.cuda.synchronize()
    L.tc32_set_debug(None)
    d = dbg.cpu()
    for cta in (0, 8, 16, 24) if M <= 1024 else (0, 5):
        r = d[cta]
        t0 = int(r[0])
        rel = lambda i: int(r[i]) - t0
        print(f"shape {(M, N, K, E)} cta {cta}: setup {rel(1)}  mainloop-done {rel(2)}  acc-ready {rel(3)}  end {rel(60)}"
              f" | plan+deps {rel(61)} staged {rel(62)} cluster-staged {rel(63)} (mma warp: {rel(58)} -> {rel(59)})")
        for i in range(6):
            print(f"   kb{i}: prod loads-issued {rel(4 + 4 * i)} stage-free {rel(5 + 4 * i)} stored {rel(6 + 4 * i)} arrived {rel(7 + 4 * i)}"
                  f" | mma landed {rel(32 + 2 * i)} issued {rel(33 + 2 * i)}")
