mkdir -p gpurun_out
for v in 1 0 1 0; do
D3B_PROLOGUE=$v timeout 300 python bench.py --steps 400 --warmup 20 --headline-only 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('prologue=$v',d['value'],d['ms_per_step'],d['e2e']['value'],d.get('graph_nodes_per_update'))"
done
