"""Achieved HBM bandwidth of the HBM-bound kernels of the path (K1 gather, K1b frame-stack gather, K10 Adam /
Adam+Polyak / soft_sync) against the measured copy peak (MEASURED_PEAKS.json hbm_gbs).  Sizes are chosen well above
the 126 MB L2 so that the traffic is real DRAM traffic; bytes are the ALGORITHMIC bytes of DESIGN.md §4.
Usage: python profiles/hbm_kernels_probe.py > profiles/r1_hbm_kernels.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3rlpy_b200._lib import lib  # noqa: E402

L, dev = lib(), torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e-3


out = {}
# ---- Adam / Adam + Polyak / soft_sync over a 64 Mi-parameter arena (256 MB per buffer)
n = 64 * 1024 * 1024
p, g, m, v, t = (torch.randn(n, device=dev) * 0.01 for _ in range(5))
step = torch.ones(1, dtype=torch.int32, device=dev)
s = timeit(lambda: L.adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), None, n, step.data_ptr(), 3e-4, 0.9,
                               0.999, 1e-8, 0.0, 1, st))
out["adam_step (28 B/param + 4 B grad zeroing)"] = {"params": n, "seconds": s, "GB/s": 32 * n / s / 1e9}
s = timeit(lambda: L.adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), t.data_ptr(), n, step.data_ptr(),
                               3e-4, 0.9, 0.999, 1e-8, 0.005, 1, st))
out["adam_step + fused soft_sync (40 B/param)"] = {"params": n, "seconds": s, "GB/s": 40 * n / s / 1e9}
s = timeit(lambda: L.soft_sync(t.data_ptr(), p.data_ptr(), n, 0.005, st))
out["soft_sync (12 B/param)"] = {"params": n, "seconds": s, "GB/s": 12 * n / s / 1e9}
del p, g, m, v, t

# ---- vector gather: c5-shaped rows (obs 111, act 8) from a 1M-step replay, batch 1M rows (read + write)
S, O, A, B = 1_000_000, 111, 8, 1_000_000
obs = torch.randn(S, O, device=dev)
act = torch.rand(S, A, device=dev)
rew = torch.randn(S, device=dev)
meta = torch.zeros(S, 4, dtype=torch.int32)
meta[:, 0] = torch.arange(S)
meta[:, 1] = (torch.arange(S) // 1000) * 1000
meta[:, 2] = meta[:, 1] + 999
meta[999::1000, 3] = 3
meta = meta.to(dev)
idx = torch.randint(0, S, (B,), device=dev, dtype=torch.int64)
o_obs, o_next = torch.empty(B, O, device=dev), torch.empty(B, O, device=dev)
o_act, o_rew, o_term, o_n = torch.empty(B, A, device=dev), torch.empty(B, device=dev), torch.empty(B, device=dev), torch.empty(B, device=dev)
s = timeit(lambda: L.gather_vector(obs.data_ptr(), O, act.data_ptr(), A, 0, rew.data_ptr(), meta.data_ptr(), idx.data_ptr(), B,
                                   1, 0.99, o_obs.data_ptr(), o_act.data_ptr(), o_rew.data_ptr(), o_next.data_ptr(),
                                   o_term.data_ptr(), o_n.data_ptr(), None, None, 0.0, st))
bytes_ = 2 * B * (2 * O + A + 3) * 4
out["gather_vector (c5 rows, 1M-row batch, read+write)"] = {"rows": B, "seconds": s, "GB/s": bytes_ / s / 1e9}
del obs, act, o_obs, o_next

# ---- frame-stack gather: c4-shaped uint8 84x84 frames, n_frames 4, batch 4096 (read + write)
S, HW, B, NF = 50_000, 84 * 84, 4096, 4
frames = torch.randint(0, 256, (S, HW), dtype=torch.uint8, device=dev)
meta = torch.zeros(S, 4, dtype=torch.int32)
meta[:, 0] = torch.arange(S)
meta[:, 1] = (torch.arange(S) // 2000) * 2000
meta[:, 2] = meta[:, 1] + 1999
meta[1999::2000, 3] = 3
meta = meta.to(dev)
idx = torch.randint(0, S, (B,), device=dev, dtype=torch.int64)
fo, fn_ = torch.empty(B, NF * HW, dtype=torch.uint8, device=dev), torch.empty(B, NF * HW, dtype=torch.uint8, device=dev)
s = timeit(lambda: L.gather_frames(frames.data_ptr(), HW, meta.data_ptr(), idx.data_ptr(), B, NF, 1, fo.data_ptr(),
                                   fn_.data_ptr(), st))
bytes_ = 2 * 2 * B * NF * HW
# the obs stack (frames g-3..g) and the next stack (g-2..g+1) of a row share n_frames-1 frames: the second read of a
# shared frame is an L2 hit, so the DRAM traffic is (n_frames+1) frame reads + 2*n_frames frame writes per row
dram = B * (NF + 1 + 2 * NF) * HW
out["gather_frames (c4 stacks, batch 4096, read+write)"] = {"rows": B, "seconds": s, "GB/s": bytes_ / s / 1e9,
                                                            "GB/s_distinct_frames_once": dram / s / 1e9}
for k in out:
    out[k]["frac_of_measured_hbm_peak"] = out[k]["GB/s"] / peak
    if "GB/s_distinct_frames_once" in out[k]:
        out[k]["frac_of_measured_hbm_peak_distinct_frames_once"] = out[k]["GB/s_distinct_frames_once"] / peak
out["_peak_GB/s"] = peak
print(json.dumps(out, indent=1))
