mkdir -p gpurun_out
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 300 python scripts/tc_trace.py "Encoder.0.layers.0.main.0,Encoder.1.layers.0.main.0,Encoder.1.layers.0.main.1" 1 1 > gpurun_out/r3e_trace.log 2>&1
unset READ_B200_LIB
grep -v "sample" gpurun_out/r3e_trace.log | cut -c1-330
grep "role 4 sample\|role  4 sample" gpurun_out/r3e_trace.log | cut -c1-400
