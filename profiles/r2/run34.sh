mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parallel.py -q -m gpu 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2_n2_b.log 2> gpurun_out/r2_n2_b.err
tail -c 300 gpurun_out/r2_n2_b.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_n2_b.log') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp_check'))
print(json.dumps(d['extra'].get('c5_strong'))[:200])
PY
