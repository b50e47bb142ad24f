mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "gather or predicate" > gpurun_out/b_t1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q > gpurun_out/b_t2.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b_smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raster|zbuf|gather_kernel|gated_conv" -c 330 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_project|gather_kernel|gated_conv" -c 13 -o gpurun_out/prof_r1 python scripts/profile_kernels.py > gpurun_out/b_ncu_full.log 2>&1
for f in b_t1 b_t2 b_smoke b_bench b_ncu_list b_ncu_full; do echo "=== $f"; tail -n 12 gpurun_out/$f.log; done
ls -la gpurun_out
