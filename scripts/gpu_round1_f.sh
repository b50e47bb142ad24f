mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/f_t1.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/layer_times_f.json > gpurun_out/f_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_project|pyramid_resolve|gated_conv" -c 8 -o gpurun_out/prof_r1f python scripts/profile_kernels.py "Encoder.0.layers.0.main.0,Encoder.1.layers.0.main.0,Encoder.2.layers.0.main.0,Encoder.3.layers.0.main.0,AFFs.0.conv.0,feat_extract.5" > gpurun_out/f_ncu_full.log 2>&1
for f in f_t1 f_smoke f_bench f_bench_ref f_ncu_list f_ncu_full; do echo "=== $f"; tail -n 5 gpurun_out/$f.log | cut -c1-3000; done
