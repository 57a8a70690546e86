mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x 2>&1 | grep -v "^$" | tail -4
timeout 300 python profiles/all_configs_bench.py > gpurun_out/r2_all_configs_b.json 2> gpurun_out/r2_all_configs_b.err
python -c "
import json
a=json.load(open('gpurun_out/r2_all_configs_b.json'))
for k,v in a.items():
    if 'cpu' not in k: print(k, round(v['us_per_update'],1))"
