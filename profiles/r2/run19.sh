mkdir -p gpurun_out
D3B_TWO_SHOT=2 timeout 600 python -m pytest tests/test_parallel.py -q -m gpu 2>&1 | tail -3
D3B_TWO_SHOT=2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "rc=$?"
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/r2_bench_n2.json') if l.startswith('{')]
d=json.loads(lines[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'nodes',d['graph_nodes_per_update'])
print('c5', d['extra']['c5_strong']['value'], d['extra']['c5_strong']['ms_per_step'])
print({k:v for k,v in d['dp_check'].items() if k!='what'})
PY
