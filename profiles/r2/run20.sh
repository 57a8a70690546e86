mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc32_gpu.py -q 2>&1 | tail -3
timeout 120 python profiles/r2/tc32_probe.py 2>&1 | grep "K=  256\|K= 4096" | tee gpurun_out/r2_tc32_probe_c.log
timeout 120 python profiles/r2/tc32_bench.py 2>&1 | tail -12 | cut -c1-140 | tee gpurun_out/r2_tc32_bench_f.log
timeout 300 python bench.py --precision fp32 --steps 200 --warmup 10 --headline-only > gpurun_out/r2_fp32_tc.json 2> gpurun_out/r2_fp32_tc.err; python -c "
import json;d=json.load(open('gpurun_out/r2_fp32_tc.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['graph_nodes_per_update'])"
