mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc32_gpu.py -q -x 2>&1 | tail -15
timeout 200 python profiles/r2/tc32_small.py 2>&1 | tee gpurun_out/r2_tc32_small_a.log | tail -20
timeout 200 python profiles/r2/tc32_phase_probe.py 2>&1 | tee gpurun_out/r2_tc32_phase_c.log | head -16
