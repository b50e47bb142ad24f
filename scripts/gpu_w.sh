mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/w_t1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/w_t2.log 2>&1
for f in w_t1 w_t2; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-400; done
