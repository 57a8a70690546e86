timeout 600 python -m pytest tests/test_update_gpu.py -q -k "checkpoint_layout" 2>&1 | grep -v "^$" | tail -30
timeout 600 python -m pytest tests/test_awr_gpu.py -q 2>&1 | tail -3
