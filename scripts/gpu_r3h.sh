mkdir -p gpurun_out
timeout 600 python scripts/bench_raster_stream.py gpurun_out/r3h_raster.json > gpurun_out/r3h_raster.log 2>&1
head -8 gpurun_out/r3h_raster.log
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_fullsize.py tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/r3h_tests.log 2>&1
echo "=== tests rc=$?"; tail -4 gpurun_out/r3h_tests.log
