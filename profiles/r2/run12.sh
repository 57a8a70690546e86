timeout 600 python -m pytest tests/test_q_function_api_gpu.py -q --tb=short 2>&1 | grep -v Warning | tail -30
timeout 900 python -m pytest tests/test_update_gpu.py -q -k "c2_shape_vs_oracle_three or bcq_c3 or c4_shape_vs" --tb=line 2>&1 | grep "Error\|passed\|failed" | tail -12
