mkdir -p gpurun_out
timeout 600 python scripts/ab_opt.py tc_pair gpurun_out/r2y_ab_pair.json 1 2 > gpurun_out/r2y_ab.log 2>&1
tail -40 gpurun_out/r2y_ab.log
