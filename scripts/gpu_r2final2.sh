mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/fin3_tests.log 2>&1
echo "=== gpu tests rc=$?"; tail -3 gpurun_out/fin3_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin3_smoke.log 2>&1
echo "=== smoke rc=$?"; tail -1 gpurun_out/fin3_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/fin3_layers.json > gpurun_out/fin3_bench.log 2> gpurun_out/fin3_bench.err
echo "=== bench rc=$?"; tail -1 gpurun_out/fin3_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
r=d['roofline']; print('roofline frac', r['frac'], 'sust', r.get('frac_of_sustained'), 'ach', r['achieved'], 'graph TF', r['whole_net_tflops_in_graph'])
print({k:(round(v['us_per_layer'],1), round(v['frac'],3)) for k,v in r['by_class'].items()})
print('raster', d['roofline_raster']['frac'], d['roofline_raster']['ms_per_frame'])
print('clocks', d['clocks']); print('parity', d['parity']['ok']); print('cpu', d['cpu_baseline']['value']); print('refgpu', d['reference_gpu'].get('tf32'), d['reference_gpu'].get('our_e2e_speedup_vs_tf32')); print('surface', d['reference_surface'])
"
tail -2 gpurun_out/fin3_bench.err
