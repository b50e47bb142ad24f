mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1
echo "=== tests rc=$?"; tail -15 gpurun_out/r2o_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu --layer-times gpurun_out/r2o_layers.json > gpurun_out/r2o_bench.log 2>&1
echo "=== bench rc=$?"; tail -3 gpurun_out/r2o_bench.log | cut -c1-1800
python scripts/show_layers.py gpurun_out/r2o_layers.json 2>&1 | head -90
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 300 python scripts/tc_trace.py "Encoder.1.layers.0.main.0,Encoder.1.layers.0.main.1" 1 1 > gpurun_out/r2o_trace_pair.log 2>&1
unset READ_B200_LIB
echo "=== trace"; grep -v "^role  *\(6\|7\|1[0-1]\|1[4-9]\):" gpurun_out/r2o_trace_pair.log | cut -c1-330 | tail -40
