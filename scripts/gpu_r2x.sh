mkdir -p gpurun_out
echo "=== a. eager SCM0.main.2 -> SCM0.main.3 under ncu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gated_conv" --csv --log-file gpurun_out/r2x_a.csv python scripts/profile_kernels.py "SCM0.main.2,SCM0.main.3,SCM0.conv" > gpurun_out/r2x_a.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2x_a.log | cut -c1-200; awk -F'","' '{print $1, $5, $NF}' gpurun_out/r2x_a.csv | cut -c1-120 | tail -4
echo "=== b. bench launch list without the streamed pair kernel"
READ_B200_OPTIONS="tc_pair_wide=0" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2x_b.csv python bench.py --steps 2 --warmup 3 --profile-timed-region > gpurun_out/r2x_b.log 2>&1
echo "rc=$? rows=$(wc -l < gpurun_out/r2x_b.csv)"; tail -2 gpurun_out/r2x_b.log | cut -c1-200
echo "=== c. bench launch list, eager replay (no graph): READ_BENCH_NO_GRAPH=1"
READ_BENCH_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2x_c.csv python bench.py --steps 2 --warmup 3 --profile-timed-region > gpurun_out/r2x_c.log 2>&1
echo "rc=$? rows=$(wc -l < gpurun_out/r2x_c.csv)"; tail -2 gpurun_out/r2x_c.log | cut -c1-200
echo "=== d. memcheck of the eager pair at C3 size"
timeout 600 compute-sanitizer --tool memcheck python scripts/profile_kernels.py "SCM0.main.2,SCM0.main.3,Encoder.2.layers.0.main.0,Encoder.3.layers.0.main.1" > gpurun_out/r2x_d.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2x_d.log | cut -c1-200
