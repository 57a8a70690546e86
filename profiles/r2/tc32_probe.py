"""Accumulation behaviour of tcgen05.mma kind::tf32 as seen through the 3xTF32 GEMM: operands that are exactly
representable in TF32 (lo = 0) isolate the accumulator's rounding; full fp32 operands show the end-to-end error.
Prints signed relative error statistics vs fp64 for growing K (all-positive operands expose a round-toward-zero
bias), next to the SIMT FFMA kernel on the same data."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200._lib import lib  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
trunc = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
g = torch.Generator().manual_seed(0)
M, N = 256, 128
for kind in ("tf32-exact positive", "tf32-exact signed", "fp32 positive", "fp32 signed"):
    for K in (32, 256, 1024, 4096):
        x = torch.rand(1, M, K, generator=g) + 0.5
        w = torch.rand(1, N, K, generator=g) + 0.5
        if "signed" in kind:
            x = x * torch.sign(torch.randn(1, M, K, generator=g))
            w = w * torch.sign(torch.randn(1, N, K, generator=g))
        if "exact" in kind:
            x, w = trunc(x), trunc(w)
        x, w = x.to(dev), w.to(dev)
        ref = torch.einsum("emk,enk->emn", x.double(), w.double())
        scale = torch.einsum("emk,enk->emn", x.double().abs(), w.double().abs())
        out = {}
        for eng in (1, 0):
            L.set_fp32_engine(eng)
            y = torch.empty(1, M, N, device=dev)
            L.linear_forward(x.data_ptr(), K, M * K, w.data_ptr(), K, N * K, None, 0, y.data_ptr(), N, M * N, M, N, K, 1, 0, st)
            torch.cuda.synchronize()
            e = (y.double() - ref) / scale
            out[eng] = (float(e.mean()), float(e.abs().max()))
        print(f"{kind:22s} K={K:5d}  tc32 mean {out[1][0]:+.2e} max {out[1][1]:.2e} | simt mean {out[0][0]:+.2e} max {out[0][1]:.2e}"
              f"   (errors relative to sum|a||b|)")
L.set_fp32_engine(1)
