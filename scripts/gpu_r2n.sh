mkdir -p gpurun_out
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 300 python scripts/tc_trace.py "Encoder.1.layers.0.main.0,Encoder.1.layers.0.main.1" 1 1 > gpurun_out/r2n_trace_pair.log 2>&1
unset READ_B200_LIB
echo "=== trace"; grep -v "^role  *\(6\|7\|1[0-1]\|1[4-9]\):" gpurun_out/r2n_trace_pair.log | cut -c1-330
