set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc32_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/r2_tc32_tests.log
cat gpurun_out/r2_tc32_tests.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "linear" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_update_gpu.py -q -k "c2_shape_vs_oracle_three or cql_matches or td3bc_c1 or bcq_c3 or c4_shape" 2>&1 | tail -15
timeout 300 python bench.py --precision fp32 --steps 200 --warmup 10 > gpurun_out/r2_fp32_tc.json 2> gpurun_out/r2_fp32_tc.err; tail -c 600 gpurun_out/r2_fp32_tc.err; python -c "
import json;d=json.load(open('gpurun_out/r2_fp32_tc.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['graph_nodes_per_update']);
[print(k,v) for k,v in d['roofline']['families'].items()]"
D3B_FP32_ENGINE=simt timeout 300 python bench.py --precision fp32 --steps 100 --warmup 10 > gpurun_out/r2_fp32_simt.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2_fp32_simt.json'));print(d['value'],d['ms_per_step'])"
