mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r2_n${N}_final.log 2> gpurun_out/r2_n${N}_final.err
tail -c 300 gpurun_out/r2_n${N}_final.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_n${N}_final.log') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp_check'))
print(json.dumps(d['extra'].get('c5_strong'))[:300])
PY
