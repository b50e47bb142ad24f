mkdir -p gpurun_out
# 1. launch list of the timed region only (2 frames of the C3 bench, CUDA-graph kernel nodes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launch_list_bench.csv python bench.py --steps 2 --warmup 3 --profile-timed-region > gpurun_out/r2v_list.log 2>&1
echo "=== launch list rc=$? rows=$(wc -l < gpurun_out/r02_launch_list_bench.csv)"; tail -2 gpurun_out/r2v_list.log | cut -c1-300
# 2. one --set full capture of each hot kernel
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"raster_|pyramid_resolve|gated_conv|upsample" -c 20 -o gpurun_out/prof_r02 python scripts/profile_kernels.py "Encoder.0.layers.0.main.0,Encoder.0.layers.0.main.1,Encoder.1.layers.0.main.0,Encoder.1.layers.0.main.1,Encoder.2.layers.0.main.0,Encoder.2.layers.0.main.1,Encoder.3.layers.0.main.0,Encoder.3.layers.0.main.1,feat_extract.1,feat_extract.7,Convs.2,AFFs.0.conv.0,AFFs.1.conv.0,feat_extract.0,feat_extract.5,FAM2.merge" > gpurun_out/r2v_full.log 2>&1
echo "=== full rc=$?"; tail -2 gpurun_out/r2v_full.log | cut -c1-300
# 3. compute-sanitizer
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_frame.py both > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "=== memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_frame.py smoke > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "=== racecheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python scripts/sanitize_frame.py smoke > gpurun_out/r02_sanitizer_synccheck.log 2>&1
echo "=== synccheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_synccheck.log
