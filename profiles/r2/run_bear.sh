set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_awr_gpu.py -q -x -k bear 2>&1 | tail -40
