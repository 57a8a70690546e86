mkdir -p gpurun_out
for i in 1 2; do
BENCH_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 100 --warmup 5 --profile-dp > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "rc=$? bytes=$(wc -c < gpurun_out/r2_bench_n2.json)"
grep "bench rank\|Error\|error" gpurun_out/r2_bench_n2.err | tail -20
done
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'nodes',d['graph_nodes_per_update'])
tot=0
for k,v in d['extra']['dp_families'].items():
    print(f"  {k:28s} {v['launches_per_update']:4.1f} launches {v['us_per_update']:8.1f} us"); tot+=v['us_per_update']
print('sum',tot)
PY
