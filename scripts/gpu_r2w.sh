mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launch_list_bench.csv python bench.py --steps 2 --warmup 3 --profile-timed-region > gpurun_out/r2w_list.log 2>&1
echo "=== launch list rc=$? rows=$(wc -l < gpurun_out/r02_launch_list_bench.csv)"; tail -2 gpurun_out/r2w_list.log | cut -c1-300
for c in c1 c2; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r02_bench_$c.json 2> gpurun_out/r2w_bench_$c.err
  echo "=== $c rc=$?"; tail -1 gpurun_out/r02_bench_$c.json | cut -c1-400
done
timeout 900 python bench.py --config c5 --steps 20 --warmup 3 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r2w_bench_c5.err
echo "=== c5 rc=$?"; tail -1 gpurun_out/r02_bench_c5.json | cut -c1-600
