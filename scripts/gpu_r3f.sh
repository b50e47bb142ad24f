mkdir -p gpurun_out
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 300 python scripts/ab_pair_dbg.py "Encoder.0.layers.0.main.0,Encoder.1.layers.0.main.0,Decoder.3.layers.1.main.0" "tc_debug=0,tc_debug=2,tc_debug=64,tc_tma_store=0" > gpurun_out/r3f_dbg.log 2>&1
unset READ_B200_LIB
tail -5 gpurun_out/r3f_dbg.log
