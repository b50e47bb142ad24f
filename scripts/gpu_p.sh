mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/p_t1.log 2>&1
timeout 300 python scripts/bench_raster_modes.py > gpurun_out/p_raster.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_times_p.json > gpurun_out/p_bench.log 2>&1
for f in p_t1 p_raster p_bench; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-1500; done
