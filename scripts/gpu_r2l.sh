mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "pair" > gpurun_out/r2l_t1.log 2>&1
timeout 600 python scripts/ab_commit.py gpurun_out/r2l_ab.json > gpurun_out/r2l_ab.log 2>&1
for f in r2l_t1 r2l_ab; do echo "=== $f"; tail -n 25 gpurun_out/$f.log | cut -c1-300; done
