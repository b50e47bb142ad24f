mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r3j_tests.log 2>&1
echo "=== tests rc=$?"; tail -3 gpurun_out/r3j_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r3j_bench.log 2> gpurun_out/r3j_bench.err
echo "=== bench rc=$?"; tail -1 gpurun_out/r3j_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
r=d['roofline']; print('roofline frac', r['frac'], 'sust', r.get('frac_of_sustained'), 'ach', r['achieved'])
print({k:(round(v['us_per_layer'],1), round(v['frac'],3)) for k,v in r['by_class'].items()})
print('raster', d['roofline_raster']['frac'], d['roofline_raster']['ms_per_frame'], d['roofline_raster']['project_ms'], d['roofline_raster']['resolve_gather_ms'])
print('clocks', d['clocks'])
"
tail -3 gpurun_out/r3j_bench.err
