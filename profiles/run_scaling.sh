#!/bin/bash
# usage: profiles/run_scaling.sh N   (inside gpurun --gpus N): c2 weak scaling and c5 strong scaling at N ranks
N=$1
run() {
  if [ "$N" = "1" ]; then timeout 300 python bench.py --gpus 1 "$@"
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29710 bench.py --gpus $N "$@"; fi
}
run --steps 200 --warmup 20 > gpurun_out/scale_c2_n$N.json 2> gpurun_out/scale_c2_n$N.err
run --workload c5 --steps 20 --warmup 5 > gpurun_out/scale_c5_n$N.json 2> gpurun_out/scale_c5_n$N.err
for f in gpurun_out/scale_c2_n$N.json gpurun_out/scale_c5_n$N.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1))
except Exception as e:
    print(sys.argv[1],'FAILED',e)
P
done
tail -3 gpurun_out/scale_c5_n$N.err
