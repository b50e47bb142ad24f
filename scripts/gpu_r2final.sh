mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/fin2_tests.log 2>&1
echo "=== gpu tests rc=$?"; tail -3 gpurun_out/fin2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin2_smoke.log 2>&1
echo "=== smoke rc=$?"; tail -1 gpurun_out/fin2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/fin2_layers.json > gpurun_out/fin2_bench.log 2> gpurun_out/fin2_bench.err
echo "=== bench rc=$?"; tail -1 gpurun_out/fin2_bench.log | cut -c1-200; tail -2 gpurun_out/fin2_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/fin2_bench_ref.log 2> gpurun_out/fin2_bench_ref.err
echo "=== ref arm rc=$?"; tail -1 gpurun_out/fin2_bench_ref.log | cut -c1-400
for c in c1 c2; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/fin2_bench_$c.json 2> gpurun_out/fin2_bench_$c.err
  echo "=== $c rc=$?"; tail -1 gpurun_out/fin2_bench_$c.json | cut -c1-160
done
timeout 900 python bench.py --config c5 --steps 20 --warmup 3 > gpurun_out/fin2_bench_c5.json 2> gpurun_out/fin2_bench_c5.err
echo "=== c5 rc=$?"; tail -1 gpurun_out/fin2_bench_c5.json | cut -c1-160
READ_BENCH_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/fin2_launch_list.csv python bench.py --steps 2 --warmup 3 --profile-timed-region > gpurun_out/fin2_list.log 2>&1
echo "=== launch list rc=$? rows=$(wc -l < gpurun_out/fin2_launch_list.csv)"
