mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1
echo "=== tests rc=$?"; tail -8 gpurun_out/r2s_tests.log
timeout 600 python scripts/ab_tma.py gpurun_out/r2s_ab_tma.json > gpurun_out/r2s_ab.log 2>&1
tail -60 gpurun_out/r2s_ab.log
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2s_tests2.log 2>&1
echo "=== tests2 rc=$?"; tail -8 gpurun_out/r2s_tests2.log
