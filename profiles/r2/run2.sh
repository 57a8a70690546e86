set -x
mkdir -p gpurun_out
timeout 120 python profiles/r2/tc32_probe.py 2>&1 | tail -20 | tee gpurun_out/r2_tc32_probe.log
timeout 300 python -m pytest tests/test_tc32_gpu.py -q 2>&1 | tail -30
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "linear" 2>&1 | tail -15
