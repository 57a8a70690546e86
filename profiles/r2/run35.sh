mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 profiles/r2/dp_trace.py c2 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee gpurun_out/r2_dp_trace_n$N.log | head -40
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 profiles/r2/dp_trace.py c5 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee -a gpurun_out/r2_dp_trace_n$N.log | head -40
