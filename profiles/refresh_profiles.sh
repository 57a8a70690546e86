#!/bin/bash
# usage (inside gpurun, 1 GPU): bash profiles/refresh_profiles.sh TAG
# plain bench first (exit 0), then the ncu launch list of the same command, then one --set full capture of the fused kernels
TAG=${1:-r1}
set -x
timeout 280 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 120 python profiles/run_c2_update.py 3 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout 280 ncu --set full --clock-control none --import-source on -k regex:mlp_ -c 12 -o gpurun_out/prof_mlp_$TAG -f \
  python profiles/run_c2_update.py 3 > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i gpurun_out/prof_mlp_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_mlp_raw_$TAG.csv 2>/dev/null
timeout 200 python profiles/all_configs_bench.py > gpurun_out/all_configs_$TAG.json 2> gpurun_out/all_configs_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json
