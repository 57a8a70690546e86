mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x 2>&1 | grep -v "^$" | tail -6
timeout 300 python bench.py --steps 300 --warmup 20 --headline-only > gpurun_out/r2_bf16_b.json 2> gpurun_out/r2_bf16_b.err
tail -c 300 gpurun_out/r2_bf16_b.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/r2_bf16_b.json') if l.startswith('{')][-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d.get('graph_nodes_per_update'))"
