mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/q_t1.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_times_q.json > gpurun_out/q_bench.log 2>&1
for f in q_t1 q_bench; do echo "=== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-400; done
