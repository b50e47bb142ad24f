mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_conv.py tests/test_gpu_gather.py -m gpu -q -x > gpurun_out/d_t1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q > gpurun_out/d_t2.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --layer-times gpurun_out/layer_times_d.json > gpurun_out/d_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_project|pyramid_resolve|gated_conv" -c 12 -o gpurun_out/prof_r1d python scripts/profile_kernels.py "Encoder.0.layers.0.main.0,Encoder.0.layers.0.main.1,Encoder.1.layers.0.main.0,Encoder.2.layers.0.main.0,AFFs.0.conv.0,feat_extract.5,Convs.2,feat_extract.0" > gpurun_out/d_ncu_full.log 2>&1
for f in d_t1 d_t2 d_smoke d_bench d_ncu_full; do echo "=== $f"; tail -n 8 gpurun_out/$f.log | cut -c1-3000; done
