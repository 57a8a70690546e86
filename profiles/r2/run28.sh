mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 10 --profile-dp > gpurun_out/r2_n2_dp.log 2> gpurun_out/r2_n2_dp.err
tail -c 400 gpurun_out/r2_n2_dp.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_n2_dp.log') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp_check'))
print(json.dumps(d['extra'].get('c5_strong'))[:300])
for k,v in d['extra']['dp_families'].items(): print(k, v)
PY
