mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | grep -v "^$" | tail -3
timeout 600 python bench.py > gpurun_out/bench_r2_n1_final.json 2> gpurun_out/bench_r2_n1_final.err || tail -c 1500 gpurun_out/bench_r2_n1_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; tail -c 600 gpurun_out/bench_r2_ref.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_n1_final.json') if l.startswith('{')][-1])
print('value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'nodes',d['graph_nodes_per_update'])
f=d['fp32_parity_mode']; print('fp32',f['value'],f['ms_per_step'],f['e2e']['value'],f['graph_nodes_per_update'])
for k,v in d['extra']['c1'].items():
    if isinstance(v,dict): print('c1',k,v['value'],v['ms_per_step'],v['e2e']['value'])
print('c5',d['extra']['c5_strong']['value'])
print('roofline',d['roofline']['frac'],d['roofline']['achieved'])
PY
