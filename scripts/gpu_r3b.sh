mkdir -p gpurun_out
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 300 python scripts/tc_trace.py "feat_extract.1,feat_extract.0,feat_extract.5,AFFs.0.conv.0,Convs.2" 1 1 > gpurun_out/r3b_trace.log 2>&1
unset READ_B200_LIB
grep -v "sample" gpurun_out/r3b_trace.log | cut -c1-250
grep "role  4 sample\|role  8 sample" gpurun_out/r3b_trace.log | cut -c1-330 | head -8
