set -x
mkdir -p gpurun_out
timeout 120 python profiles/r2/tc32_probe.py 2>&1 | tail -20 | tee gpurun_out/r2_tc32_probe.log
timeout 300 python -m pytest tests/test_tc32_gpu.py -q 2>&1 | tail -30
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "linear" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_update_gpu.py -q -x 2>&1 | tail -15
timeout 300 python bench.py --precision fp32 --steps 200 --warmup 10 > gpurun_out/r2_fp32_tc.json 2> gpurun_out/r2_fp32_tc.err; tail -c 600 gpurun_out/r2_fp32_tc.err; python -c "
import json;d=json.load(open('gpurun_out/r2_fp32_tc.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['graph_nodes_per_update']);
[print(k,v) for k,v in d['roofline']['families'].items() if k.startswith('linear')]"
