mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "pair" > gpurun_out/r3d_tests.log 2>&1
echo "=== pair tests rc=$?"; tail -4 gpurun_out/r3d_tests.log
timeout 600 python scripts/ab_opt.py tc_pair_wide gpurun_out/r3d_ab_wide.json > gpurun_out/r3d_ab.log 2>&1
grep -v "layers\.[123]\." gpurun_out/r3d_ab.log | tail -24
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r3d_tests2.log 2>&1
echo "=== tests2 rc=$?"; tail -3 gpurun_out/r3d_tests2.log
