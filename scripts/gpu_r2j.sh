mkdir -p gpurun_out
timeout 300 python scripts/debug_rmsprop.py > gpurun_out/r2j_dbg.log 2>&1
timeout 300 python scripts/bench_raster_stream.py gpurun_out/r2j_raster.json > gpurun_out/r2j_raster.log 2>&1
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/r2j_t1.log 2>&1
for f in r2j_dbg r2j_raster r2j_t1; do echo "=== $f"; tail -n 14 gpurun_out/$f.log | cut -c1-400; done
