# round 2, call A: new full-size parity tests + baseline bench of the round-1 kernels on this box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_conv.py -m gpu -q -x -k "fullsize or c3 or c2 or persistent" > gpurun_out/r2a_t1.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --layer-times gpurun_out/layer_times_r2a.json --no-cpu-baseline > gpurun_out/r2a_bench.log 2>&1
for f in r2a_t1 r2a_bench; do echo "=== $f"; tail -n 6 gpurun_out/$f.log | cut -c1-1500; done
