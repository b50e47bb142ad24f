mkdir -p gpurun_out
timeout 600 python scripts/ab_side.py 16,24,32 > gpurun_out/r3g_side.log 2>&1
tail -4 gpurun_out/r3g_side.log
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_fullsize.py tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/r3g_tests.log 2>&1
echo "=== tests rc=$?"; tail -4 gpurun_out/r3g_tests.log
