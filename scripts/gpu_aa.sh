mkdir -p gpurun_out
timeout 400 python scripts/bench_raster_modes.py > gpurun_out/aa_raster.log 2>&1
tail -n 8 gpurun_out/aa_raster.log
