mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/x_t1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/x_t2.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_times_x.json > gpurun_out/x_bench.log 2>&1
for f in x_t1 x_t2; do echo "=== $f"; tail -n 14 gpurun_out/$f.log | cut -c1-500; done
python scripts/show_layers.py gpurun_out/layer_times_x.json 0.0 | grep "AFFs\|total"
python -c "
import json
for l in open('gpurun_out/x_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('FPS', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['achieved'], d['config']['conv_impl'])
"
tail -3 gpurun_out/x_bench.log | cut -c1-600
