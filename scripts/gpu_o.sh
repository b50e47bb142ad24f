mkdir -p gpuruo_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -q -x > gpuruo_out/o_t1.log 2>&1
timeout 300 python scripts/tc_debug_times.py > gpuruo_out/o_dbg.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-times gpuruo_out/layer_times_o.json > gpuruo_out/o_bench.log 2>&1
for f in o_t1 o_dbg; do echo "=== $f"; tail -n 22 gpuruo_out/$f.log | cut -c1-600; done
python scripts/show_layers.py gpuruo_out/layer_times_o.json 0.06
