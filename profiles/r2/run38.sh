mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 profiles/r2/dp_trace.py c5 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|Warning\|ret = \|out.append" | tee gpurun_out/r2_dp_trace_c5_n$N.log | head -90
