mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc32_gpu.py -q 2>&1 | tail -4
timeout 120 python profiles/r2/tc32_bench.py 2>&1 | tail -12 | tee gpurun_out/r2_tc32_bench_c.log
timeout 120 python profiles/r2/tc32_variants.py 2>&1 | tail -12 | tee gpurun_out/r2_tc32_variants_c.log
