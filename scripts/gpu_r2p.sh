mkdir -p gpurun_out
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 300 python scripts/ab_pair_dbg.py > gpurun_out/r2p_pair_dbg.log 2>&1
unset READ_B200_LIB
cat gpurun_out/r2p_pair_dbg.log | tail -20
