mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 40 --warmup 5 --workload c5 --headline-only > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err
echo "rc=$?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
lines=[l for l in open(f'gpurun_out/r2_bench_n{N}.json') if l.startswith('{')]
d=json.loads(lines[-1])
print('c2 weak value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'nodes',d['graph_nodes_per_update'])
print('c5', d['extra']['c5_strong']['value'], d['extra']['c5_strong']['ms_per_step'])
print({k:v for k,v in d['dp_check'].items() if k!='what'})
lines=[l for l in open(f'gpurun_out/r2_bench_c5_n{N}.json') if l.startswith('{')]
d=json.loads(lines[-1])
print('c5 strong value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['step_ms_p10_p50_p90'])
PY
