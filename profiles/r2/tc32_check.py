"""Value check of the dispatching entry points (linear_forward / backward_data / backward_weight, engine = 3xTF32) at
the c2 layer shapes against fp64: worst |err| / (sum |a||b|)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3rlpy_b200._lib import lib  # noqa: E402

L = lib()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(0)
for (M, N, K, E, shared) in [(7936, 256, 23, 2, True), (7936, 256, 256, 2, False), (256, 256, 23, 2, True), (512, 256, 17, 1, True),
                             (256, 256, 256, 2, False), (7936, 256, 23, 2, False)]:
    x = torch.randn(1 if shared else E, M, K, generator=g).to(dev)
    w = (torch.randn(E, N, K, generator=g) / K ** 0.5).to(dev)
    dy = torch.randn(E, M, N, generator=g).to(dev)
    xe = x.expand(E, M, K).double()
    sx = 0 if shared else M * K
    for rep in range(3):
        dw = torch.zeros(E, N, K, device=dev)
        db = torch.zeros(E, N, device=dev)
        L.linear_backward_weight(dy.data_ptr(), N, M * N, x.data_ptr(), K, sx, dw.data_ptr(), K, N * K, db.data_ptr(), N, M, N, K, E, st)
        torch.cuda.synchronize()
        ref = torch.einsum("emn,emk->enk", dy.double(), xe)
        bound = torch.einsum("emn,emk->enk", dy.double().abs(), xe.abs())
        e1 = float(((dw.double() - ref).abs() / bound).max())
        e2 = float(((db.double() - dy.double().sum(1)).abs() / dy.double().abs().sum(1)).max())
        dx = torch.empty(E, M, K, device=dev)
        L.linear_backward_data(dy.data_ptr(), N, M * N, w.data_ptr(), K, N * K, dx.data_ptr(), K, M * K, None, 0, 0, M, N, K, E, st)
        torch.cuda.synchronize()
        refx = torch.einsum("emn,enk->emk", dy.double(), w.double())
        bx = torch.einsum("emn,enk->emk", dy.double().abs(), w.double().abs())
        e3 = float(((dx.double() - refx).abs() / bx).max())
        y = torch.empty(E, M, N, device=dev)
        L.linear_forward(x.data_ptr(), K, sx, w.data_ptr(), K, N * K, None, 0, y.data_ptr(), N, M * N, M, N, K, E, 0, st)
        torch.cuda.synchronize()
        refy = torch.einsum("emk,enk->emn", xe, w.double())
        by = torch.einsum("emk,enk->emn", xe.abs(), w.double().abs())
        e4 = float(((y.double() - refy).abs() / by).max())
        print(f"{(M, N, K, E, shared)} rep {rep}: wgrad {e1:.2e} bias {e2:.2e} dgrad {e3:.2e} fwd {e4:.2e}")
