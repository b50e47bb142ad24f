mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/y_t2.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_times_y.json > gpurun_out/y_bench.log 2>&1
tail -n 3 gpurun_out/y_t2.log
python scripts/show_layers.py gpurun_out/layer_times_y.json 0.0 | grep "AFFs\|total"
python -c "
import json
for l in open('gpurun_out/y_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('FPS', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['kernel'], d['config']['conv_impl']); print(d['breakdown_ms_per_frame'])
"
