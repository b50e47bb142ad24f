mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "reversed_tile_order" > gpurun_out/r3r_tests.log 2>&1
echo "=== tests rc=$?"; tail -6 gpurun_out/r3r_tests.log
