mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1
echo "=== tests rc=$?"; tail -8 gpurun_out/r2q_tests.log
timeout 300 python scripts/ab_pair_dbg.py "Encoder.1.layers.0.main.0,Encoder.1.layers.0.main.1,Decoder.1.layers.1.main.1,Decoder.2.layers.0.main.0" "tc_tma_store=0,tc_tma_store=1" > gpurun_out/r2q_ab.log 2>&1
tail -6 gpurun_out/r2q_ab.log
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2q_tests2.log 2>&1
echo "=== tests2 rc=$?"; tail -8 gpurun_out/r2q_tests2.log
