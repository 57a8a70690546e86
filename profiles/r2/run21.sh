mkdir -p gpurun_out
timeout 200 python profiles/r2/tc32_phase_probe.py 2>&1 | tee gpurun_out/r2_tc32_phase_b.log | tail -60
