mkdir -p gpurun_out
timeout 900 python scripts/ab_commit.py gpurun_out/r2f_ab.json > gpurun_out/r2f_ab.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_stream|pyramid_resolve" -c 4 -o gpurun_out/prof_r2_raster python scripts/profile_raster.py > gpurun_out/r2f_ncu.log 2>&1
for f in r2f_ab r2f_ncu; do echo "=== $f"; tail -n 22 gpurun_out/$f.log | cut -c1-300; done
