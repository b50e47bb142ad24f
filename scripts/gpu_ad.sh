mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/ad_t1.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ad_smoke.log 2>&1
timeout 400 python scripts/bench_raster_modes.py > gpurun_out/ad_raster.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/layer_times_ad.json > gpurun_out/ad_bench.log 2>&1
for f in ad_t1 ad_smoke; do echo "=== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-600; done
tail -n 6 gpurun_out/ad_raster.log
python -c "
import json
for l in open('gpurun_out/ad_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('FPS', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac']); print(d['roofline_raster']); print(d['cpu_baseline'])
"
