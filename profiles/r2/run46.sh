timeout 900 python -m pytest tests/test_update_gpu.py tests/test_hooks_gpu.py tests/test_awr_gpu.py -q -k "bcq or plas or hooks" 2>&1 | tail -3
timeout 300 python - <<'PY'
import json, bench
d = bench.other_configs_e2e("bf16")
print({k: round(v["us_per_update"], 1) for k, v in d.items() if isinstance(v, dict)})
PY
