mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/fin_t1.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/layer_times_fin.json > gpurun_out/fin_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin_bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/fin_ncu_list.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"raster_|pyramid_resolve|gated_conv|upsample" -c 18 -o gpurun_out/prof_r1fin python scripts/profile_kernels.py "Encoder.0.layers.0.main.0,Encoder.0.layers.0.main.1,Encoder.1.layers.0.main.0,Encoder.1.layers.0.main.1,Encoder.2.layers.0.main.0,Encoder.2.layers.0.main.1,Encoder.3.layers.0.main.0,Encoder.3.layers.0.main.1,feat_extract.1,feat_extract.7,Convs.2,AFFs.0.conv.0,AFFs.1.conv.0,feat_extract.0,feat_extract.5" > gpurun_out/fin_ncu_full.log 2>&1
for f in fin_t1 fin_smoke fin_bench fin_bench_ref fin_ncu_full; do echo "=== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-700; done
