timeout 900 python -m pytest tests/test_update_gpu.py tests/test_hooks_gpu.py tests/test_scalers_gpu.py tests/test_online_gpu.py -q -k "td3 or ddpg or sibling or scal or online" 2>&1 | tail -3
timeout 300 python - <<'PY'
import bench, json, torch
for prec in ("bf16", "fp32"):
    r = bench.Runner(bench.WORKLOADS["c1"], 1, 0, 0, prec, False)
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    t, s, l, n = r.timed(300, 20, flush)
    print(prec, n, bench.summarize(r, t, s, 300))
PY
