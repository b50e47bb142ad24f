mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1
echo "=== gpu tests rc=$?"; tail -6 gpurun_out/r2u_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/r2u_layers.json > gpurun_out/r2u_bench.log 2>&1
echo "=== bench rc=$?"; tail -1 gpurun_out/r2u_bench.log
