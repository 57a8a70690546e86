mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc32_gpu.py -q 2>&1 | tail -8
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "linear" 2>&1 | tail -3
timeout 120 python profiles/r2/tc32_probe.py 2>&1 | grep "K=  256\|K= 4096" | tee gpurun_out/r2_tc32_probe_b.log
timeout 120 python profiles/r2/tc32_bench.py 2>&1 | tail -12 | tee gpurun_out/r2_tc32_bench_b.log
timeout 120 python profiles/r2/tc32_variants.py 2>&1 | tail -12 | tee gpurun_out/r2_tc32_variants_b.log
timeout 120 python profiles/r2/tc32_phase_probe.py 2>&1 | tail -14 | tee gpurun_out/r2_tc32_phase_b.log
