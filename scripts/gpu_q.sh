mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/q_t1.log 2>&1
timeout 300 python scripts/tc_debug_times.py "Convs.2,AFFs.0.conv.0,feat_extract.0,feat_extract.1,AFFs.1.conv.0" > gpurun_out/q_dbg.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_times_q.json > gpurun_out/q_bench.log 2>&1
for f in q_t1 q_dbg; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-600; done
python scripts/show_layers.py gpurun_out/layer_times_q.json 0.055
