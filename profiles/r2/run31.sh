mkdir -p gpurun_out
for c in c3 c4; do
timeout 120 python profiles/run_${c}_update.py > gpurun_out/plain_$c.log 2>&1 || { tail -5 gpurun_out/plain_$c.log; continue; }
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_${c}_launches.csv \
  python profiles/run_${c}_update.py > gpurun_out/ncu_$c.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_${c}_launches.csv | head -24
done
