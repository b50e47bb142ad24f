mkdir -p gpurun_out
timeout 900 python scripts/ab_layers.py gpurun_out/r2d_ab.json > gpurun_out/r2d_ab.log 2>&1
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 600 python scripts/tc_trace.py "Encoder.0.layers.0.main.0,Encoder.0.layers.0.main.1,Encoder.1.layers.0.main.0" 1 > gpurun_out/r2d_trace_mt1.log 2>&1
timeout 600 python scripts/tc_trace.py "Encoder.0.layers.0.main.0,Encoder.0.layers.0.main.1" 4 > gpurun_out/r2d_trace_mt4.log 2>&1
unset READ_B200_LIB
timeout 1200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2d_t1.log 2>&1
for f in r2d_ab r2d_t1; do echo "=== $f"; tail -n 30 gpurun_out/$f.log | cut -c1-300; done
for f in r2d_trace_mt1 r2d_trace_mt4; do echo "=== $f"; grep -v "^role  *\(5\|6\|7\|8\|9\|1[0-9]\):" gpurun_out/$f.log | grep -v sample | cut -c1-250; done
