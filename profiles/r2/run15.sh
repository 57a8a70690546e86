mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "rc=$?"
wc -c gpurun_out/r2_bench_n2.json gpurun_out/r2_bench_n2.err
grep -v "Warning\|warn\|^\*\|OMP_NUM" gpurun_out/r2_bench_n2.err | tail -30
head -c 3000 gpurun_out/r2_bench_n2.json
