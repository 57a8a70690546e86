mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err ) 2>&1 | grep real
tail -c 2500 gpurun_out/r2_bench_n2.err | grep -v Warning
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'nodes',d['graph_nodes_per_update'],d['config']['parallelism'])
print('c5',d['extra']['c5_strong'])
print('check',d['dp_check'])
PY
timeout 600 python -m pytest tests/test_parallel.py -q -m gpu 2>&1 | tail -3
