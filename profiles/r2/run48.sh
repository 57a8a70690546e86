mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_update_gpu.py tests/test_hooks_gpu.py -q -k "cql or sac" 2>&1 | tail -2
for v in 1 0 1 0; do
for prec in bf16 fp32; do
D3B_ALPHA_BRANCH=$v timeout 300 python bench.py --steps 300 --warmup 20 --headline-only --precision $prec 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('alpha_branch=$v $prec',round(d['value'],1),round(d['ms_per_step']*1e3,2),round(d['e2e']['value'],1),d.get('graph_nodes_per_update'))"
done
done
