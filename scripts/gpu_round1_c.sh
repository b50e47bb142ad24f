mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/c_t1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_gather.py -m gpu -q > gpurun_out/c_t2.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_times_c.json > gpurun_out/c_bench.log 2>&1
for f in c_t1 c_t2 c_bench; do echo "=== $f"; tail -n 15 gpurun_out/$f.log; done
