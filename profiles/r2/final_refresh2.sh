#!/bin/bash
# usage (inside gpurun, 1 GPU): bash profiles/r2/final_refresh2.sh — ncu evidence of the final round-2 code
# every ncu pass follows a plain run of the same command that exited 0; numbers printed under ncu are never bench values
set -x
mkdir -p gpurun_out
timeout 280 python bench.py --steps 2 --warmup 3 --headline-only > gpurun_out/plain_bf16.log 2>&1 &&
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bf16_launches.csv \
  python bench.py --steps 2 --warmup 3 --headline-only > gpurun_out/ncu_launches_bf16.log 2>&1
timeout 280 python bench.py --steps 2 --warmup 3 --headline-only --precision fp32 > gpurun_out/plain_fp32.log 2>&1 &&
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_fp32_launches.csv \
  python bench.py --steps 2 --warmup 3 --headline-only --precision fp32 > gpurun_out/ncu_launches_fp32.log 2>&1
timeout 120 python profiles/run_c2_update.py 3 > gpurun_out/plain_bf16_update.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_ -s 14 -c 14 -o gpurun_out/prof_mlp_r2 -f \
  python profiles/run_c2_update.py 3 > gpurun_out/ncu_full_mlp.log 2>&1
ncu -i gpurun_out/prof_mlp_r2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_mlp_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -8
