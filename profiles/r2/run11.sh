mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --tb=line 2>&1 | grep -v Warning | tail -14
