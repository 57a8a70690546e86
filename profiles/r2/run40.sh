timeout 600 python -m pytest tests/test_awr_gpu.py -q -k "round_trip" 2>&1 | grep -v "^$" | tail -40
