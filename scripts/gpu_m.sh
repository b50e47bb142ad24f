mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/m_t1.log 2>&1
timeout 300 python scripts/tc_debug_times.py > gpurun_out/m_dbg.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_times_m.json > gpurun_out/m_bench.log 2>&1
for f in m_t1 m_dbg; do echo "=== $f"; tail -n 22 gpurun_out/$f.log | cut -c1-600; done
python scripts/show_layers.py gpurun_out/layer_times_m.json 0.06
