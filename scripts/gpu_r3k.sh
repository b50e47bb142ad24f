mkdir -p gpurun_out
timeout 600 python scripts/ab_frame_opt.py "raster_carveout=-1,raster_stages=3;raster_carveout=45,raster_stages=2;raster_carveout=65,raster_stages=3;raster_carveout=100,raster_stages=3;raster_stream=0;raster_stream=1,raster_carveout=45,raster_stages=2" > gpurun_out/r3k_ab.log 2>&1
tail -7 gpurun_out/r3k_ab.log
