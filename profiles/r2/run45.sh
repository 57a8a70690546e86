timeout 900 python -m pytest tests/test_update_gpu.py tests/test_hooks_gpu.py tests/test_online_gpu.py tests/test_q_function_api_gpu.py -q -k "dqn or discrete or qr or nfq or dcql or online or pixel" 2>&1 | tail -3
timeout 300 python - <<'PY'
import json, bench
print(json.dumps(bench.other_configs_e2e("bf16"), indent=0))
PY
