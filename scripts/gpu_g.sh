mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/g_t1.log 2>&1
timeout 300 python scripts/bench_raster_modes.py > gpurun_out/g_raster.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_times_g.json > gpurun_out/g_bench.log 2>&1
for f in g_t1 g_raster g_bench; do echo "=== $f"; tail -n 22 gpurun_out/$f.log | cut -c1-600; done
