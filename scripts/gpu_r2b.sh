mkdir -p gpurun_out
timeout 900 python scripts/ab_conv.py gpurun_out/r2b_ab.json > gpurun_out/r2b_ab.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/r2b_t1.log 2>&1
for f in r2b_ab r2b_t1; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-900; done
