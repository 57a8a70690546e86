mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc32_gpu.py -q -x 2>&1 | tail -3 || exit 1
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k linear 2>&1 | tail -3
timeout 600 python -m pytest tests/test_update_gpu.py -q -x -k "cql_matches or td3bc_matches" --tb=line 2>&1 | tail -4
