mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/t_t1.log 2>&1
timeout 300 python scripts/tc_debug_times.py "Encoder.0.layers.0.main.0,Encoder.0.layers.0.main.1,Encoder.1.layers.0.main.0,Encoder.1.layers.0.main.1,Encoder.2.layers.0.main.0,Encoder.3.layers.0.main.0" > gpurun_out/t_dbg.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_timet_s.json > gpurun_out/t_bench.log 2>&1
for f in t_t1 t_dbg; do echo "=== $f"; tail -n 12 gpurun_out/$f.log | cut -c1-600; done
python scripts/show_layers.py gpurun_out/layer_timet_s.json 0.055
python -c "
import json
for l in open('gpurun_out/t_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('FPS', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['achieved'])
"
