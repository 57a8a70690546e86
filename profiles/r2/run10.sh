mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=line -x 2>&1 | grep -v Warning | tail -12
timeout 300 python bench.py --precision fp32 --steps 200 --warmup 10 > gpurun_out/r2_fp32_tc.json 2> gpurun_out/r2_fp32_tc.err; tail -c 400 gpurun_out/r2_fp32_tc.err; python -c "
import json;d=json.load(open('gpurun_out/r2_fp32_tc.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['graph_nodes_per_update']);
[print(k,v) for k,v in d['roofline']['families'].items() if 'linear' in k or 'head' in k]"
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/r2_bf16_a.json 2> gpurun_out/r2_bf16_a.err; tail -c 400 gpurun_out/r2_bf16_a.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bf16_a.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['graph_nodes_per_update'])"
