mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_update_gpu.py tests/test_kernels_gpu.py -q -x 2>&1 | tail -5
timeout 300 python bench.py --precision fp32 --steps 200 --warmup 10 --headline-only > gpurun_out/r2_fp32_b.json 2> gpurun_out/r2_fp32_b.err
tail -c 300 gpurun_out/r2_fp32_b.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/r2_fp32_b.json') if l.startswith('{')][-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('graph_nodes_per_update'))"
