mkdir -p gpurun_out
timeout 600 python scripts/bench_raster_stream.py gpurun_out/r3i_raster.json > gpurun_out/r3i_raster.log 2>&1
head -7 gpurun_out/r3i_raster.log
