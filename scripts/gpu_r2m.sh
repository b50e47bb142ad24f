mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_t1.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/layer_times_r2m.json > gpurun_out/r2m_bench.log 2> gpurun_out/r2m_bench.err
for f in r2m_t1 r2m_smoke; do echo "=== $f"; tail -n 8 gpurun_out/$f.log | cut -c1-300; done
for f in r2m_bench; do echo "=== $f"; tail -n 2 gpurun_out/$f.log | cut -c1-1500; tail -n 5 gpurun_out/$f.err | cut -c1-400; done
