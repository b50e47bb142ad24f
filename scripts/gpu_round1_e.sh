mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/e_t1.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/e_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/layer_times_e.json > gpurun_out/e_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/e_bench_ref.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_project|pyramid_resolve|gated_conv" -c 8 -o gpurun_out/prof_r1e python scripts/profile_kernels.py "Encoder.0.layers.0.main.0,Encoder.1.layers.0.main.0,Encoder.2.layers.0.main.0,Encoder.3.layers.0.main.0,AFFs.0.conv.0,feat_extract.5" > gpurun_out/e_ncu_full.log 2>&1
for f in e_t1 e_smoke e_bench e_bench_ref e_ncu_full; do echo "=== $f"; tail -n 5 gpurun_out/$f.log | cut -c1-3000; done
