mkdir -p gpurun_out
for eng in simt tc; do
  echo "== engine $eng"
  D3B_FP32_ENGINE=$eng timeout 300 python profiles/grad_parity_probe.py 2>&1 | grep "fp32" | tee gpurun_out/r2_grad_parity_$eng.log
done
