mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/u_t1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/u_t2.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_times_u.json > gpurun_out/u_bench.log 2>&1
for f in u_t1 u_t2; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-400; done
python scripts/show_layers.py gpurun_out/layer_times_u.json 0.03 | grep "feat_extract\|total"
python -c "
import json
for l in open('gpurun_out/u_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('FPS', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['achieved'], d['config']['conv_impl'])
"
