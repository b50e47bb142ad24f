mkdir -p gpurun_out
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 600 python scripts/tc_trace.py "Encoder.0.layers.0.main.0,Encoder.0.layers.0.main.1,Encoder.1.layers.0.main.0,Encoder.1.layers.0.main.1" 1 > gpurun_out/r2c_trace_mt1.log 2>&1
timeout 600 python scripts/tc_trace.py "Encoder.0.layers.0.main.0,Encoder.0.layers.0.main.1" 4 > gpurun_out/r2c_trace_mt4.log 2>&1
unset READ_B200_LIB
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_fullsize.py -m gpu -q -x -k "sorted or c3_raster" > gpurun_out/r2c_t1.log 2>&1
timeout 600 python scripts/bench_raster_stream.py gpurun_out/r2c_raster.json > gpurun_out/r2c_raster.log 2>&1
for f in r2c_trace_mt1 r2c_trace_mt4 r2c_t1 r2c_raster; do echo "=== $f"; tail -n 40 gpurun_out/$f.log | cut -c1-400; done
