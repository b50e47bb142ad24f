mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r3q_bench.log 2> gpurun_out/r3q_bench.err
echo "=== bench rc=$?"; tail -1 gpurun_out/r3q_bench.log | cut -c1-180; tail -2 gpurun_out/r3q_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
