mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_umma_gpu.py -q -x 2>&1 | tail -6
