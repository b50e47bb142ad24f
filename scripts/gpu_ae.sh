mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raster.py -m gpu -q -x -k sorted > gpurun_out/ae_t1.log 2>&1
timeout 400 python scripts/bench_raster_modes.py > gpurun_out/ae_raster.log 2>&1
tail -n 3 gpurun_out/ae_t1.log | cut -c1-300; tail -n 7 gpurun_out/ae_raster.log
