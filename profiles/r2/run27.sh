mkdir -p gpurun_out
timeout 120 python profiles/run_c2_update.py 6 fp32 > gpurun_out/plain_fp32_update.log 2>&1 &&
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_fp32_launches_b.csv \
  python profiles/run_c2_update.py 6 fp32 > gpurun_out/ncu_launches_fp32.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_fp32_launches_b.csv > gpurun_out/r2_fp32_launches_b.md 2>&1
head -40 gpurun_out/r2_fp32_launches_b.md
