mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | grep real
tail -c 1500 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'nodes',d['graph_nodes_per_update'],'launches',d['gpu_launches'])
print('fp32',{k:(v if not isinstance(v,dict) else v.get('value')) for k,v in d['fp32_parity_mode'].items() if k in('value','ms_per_step','e2e','graph_nodes_per_update')})
print('c1',{p:(d['extra']['c1'][p]['value'],d['extra']['c1'][p]['ms_per_step'],d['extra']['c1'][p]['e2e']['value']) for p in ('bf16','fp32')})
print('c5',d['extra']['c5_strong']['value'],d['extra']['c5_strong']['ms_per_step'])
print('roof',d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['dominant_launch'])
for k,v in d['hbm'].items():
    if isinstance(v,dict): print('  hbm',k,round(v['us'],2),'us',round(v['gbs'],1),'GB/s',round(v['frac'],3))
print('cpu',d['cpu_baseline'])
print('variant',d['variant_alpha_lr0'])
PY
