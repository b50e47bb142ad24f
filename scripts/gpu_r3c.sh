mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r3c_tests.log 2>&1
echo "=== tests rc=$?"; tail -4 gpurun_out/r3c_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/r3c_layers.json > gpurun_out/r3c_bench.log 2> gpurun_out/r3c_bench.err
echo "=== bench rc=$?"; tail -1 gpurun_out/r3c_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
r=d['roofline']; print('roofline frac', r['frac'], 'sust', r.get('frac_of_sustained'), 'ach', r['achieved']); print(json.dumps(r['by_class'], indent=0)[:1200])
print('raster', d['roofline_raster']['frac'], d['roofline_raster']['ms_per_frame'])
print('parity', d['parity']['ok'], d['parity'].get('vs_cpu_oracle'))
print('surface', d['reference_surface']); print('refgpu', d['reference_gpu'].get('tf32'), d['reference_gpu'].get('our_e2e_speedup_vs_tf32'))
print('breakdown', d['breakdown_ms_per_frame'])
"
tail -3 gpurun_out/r3c_bench.err
python scripts/show_layers.py gpurun_out/r3c_layers.json 0.08
