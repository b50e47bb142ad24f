mkdir -p gpurun_out
timeout 900 python scripts/ab_commit.py gpurun_out/r2e_ab.json > gpurun_out/r2e_ab.log 2>&1
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 300 python scripts/raster_breakdown.py > gpurun_out/r2e_raster.log 2>&1
unset READ_B200_LIB
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_raster.py -m gpu -q -x > gpurun_out/r2e_t1.log 2>&1
for f in r2e_ab r2e_raster r2e_t1; do echo "=== $f"; tail -n 22 gpurun_out/$f.log | cut -c1-300; done
