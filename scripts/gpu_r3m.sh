mkdir -p gpurun_out
timeout 600 python scripts/ab_env.py READ_B200_ALT_ORDER 0 1 > gpurun_out/r3m_ab.log 2>&1
tail -3 gpurun_out/r3m_ab.log
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r3m_tests.log 2>&1
echo "=== tests rc=$?"; tail -3 gpurun_out/r3m_tests.log
