#!/bin/bash
# usage (inside gpurun, 1 GPU): bash profiles/r2/run_s2_refresh.sh — launch list of the final code (bf16 mode)
# the ncu pass follows a plain run of the same command that exited 0; numbers printed under ncu are never bench values
mkdir -p gpurun_out
timeout 100 python bench.py --steps 2 --warmup 3 --headline-only > gpurun_out/s2_plain_bf16.log 2>&1 &&
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s2_bf16_launches.csv \
  python bench.py --steps 2 --warmup 3 --headline-only > gpurun_out/s2_ncu_launches_bf16.log 2>&1
echo ncu rc=$?
wc -l gpurun_out/s2_bf16_launches.csv
