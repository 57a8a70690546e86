mkdir -p gpurun_out
timeout 120 python profiles/run_c4_update.py fp32 > gpurun_out/plain_c4f.log 2>&1 || { tail -5 gpurun_out/plain_c4f.log; exit 1; }
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_c4_fp32_launches.csv \
  python profiles/run_c4_update.py fp32 > gpurun_out/ncu_c4f.log 2>&1
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/r2_c4_fp32_launches.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
idx=[i for i,r in enumerate(rows) if 'tick_kernel' in r['Kernel Name']]
s=idx[-1]; tot=0
for r in rows[s:]:
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']; v={'ns':v/1e3,'us':v,'ms':v*1e3}.get(u,v); tot+=v
    print(f"{re.sub(r'\(.*','',r['Kernel Name'])[:50]:50s} grid {r['Grid Size']:16s} {v:7.1f}")
print(tot)
PY
