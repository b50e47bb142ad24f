mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_gather.py tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/r2i_t1.log 2>&1
timeout 300 python scripts/bench_raster_stream.py gpurun_out/r2i_raster.json > gpurun_out/r2i_raster.log 2>&1
timeout 600 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/r2i_bench_c5.log 2> gpurun_out/r2i_bench_c5.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu --layer-times gpurun_out/layer_times_r2i.json > gpurun_out/r2i_bench.log 2> gpurun_out/r2i_bench.err
for f in r2i_t1 r2i_raster; do echo "=== $f"; tail -n 14 gpurun_out/$f.log | cut -c1-300; done
for f in r2i_bench_c5 r2i_bench; do echo "=== $f"; tail -n 2 gpurun_out/$f.log | cut -c1-3500; tail -n 8 gpurun_out/$f.err | cut -c1-400; done
