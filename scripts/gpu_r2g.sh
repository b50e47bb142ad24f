mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_t1.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
timeout 300 python scripts/bench_raster_stream.py gpurun_out/r2g_raster.json > gpurun_out/r2g_raster.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/layer_times_r2g.json > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench.err
timeout 300 python bench.py --config c2 --steps 20 --warmup 3 > gpurun_out/r2g_bench_c2.log 2> gpurun_out/r2g_bench_c2.err
timeout 300 python bench.py --config c1 --steps 20 --warmup 3 > gpurun_out/r2g_bench_c1.log 2> gpurun_out/r2g_bench_c1.err
for f in r2g_t1 r2g_smoke r2g_raster; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-300; done
for f in r2g_bench r2g_bench_c2 r2g_bench_c1; do echo "=== $f"; tail -n 2 gpurun_out/$f.log | cut -c1-6000; tail -n 5 gpurun_out/$f.err | cut -c1-400; done
