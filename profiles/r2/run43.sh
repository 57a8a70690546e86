mkdir -p gpurun_out
SECONDS=0; timeout 600 python bench.py > gpurun_out/bench_r2_n1_final.json 2> gpurun_out/bench_r2_n1_final.err || tail -c 1500 gpurun_out/bench_r2_n1_final.err
echo "bench wall seconds: $SECONDS"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_n1_final.json') if l.startswith('{')][-1])
print('value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'])
print(json.dumps(d['extra']['c3_c4_e2e'],indent=0))
PY
