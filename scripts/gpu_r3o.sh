mkdir -p gpurun_out
timeout 600 python scripts/ab_env.py opt:tc_pair 1 2 0 > gpurun_out/r3o_ab.log 2>&1
tail -3 gpurun_out/r3o_ab.log
