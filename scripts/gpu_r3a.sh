mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/r3a_tests.log 2>&1
echo "=== tests rc=$?"; tail -3 gpurun_out/r3a_tests.log
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 600 python scripts/ab_pair_dbg.py "feat_extract.1,feat_extract.2,feat_extract.3,feat_extract.0,Convs.2,feat_extract.5,AFFs.0.conv.0,AFFs.0.conv.1,SCM2.main.2,FAM2.merge" "tc_debug=0,tc_debug=2,tc_debug=4,tc_debug=8,tc_debug=6" > gpurun_out/r3a_dbg.log 2>&1
unset READ_B200_LIB
cat gpurun_out/r3a_dbg.log | tail -14
