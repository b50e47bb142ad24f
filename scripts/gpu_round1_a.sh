mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_gather.py tests/test_gpu_reference_kernel.py -m gpu -q > gpurun_out/t1.log 2>&1
timeout 400 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "generic" > gpurun_out/t2.log 2>&1
timeout 400 python -m pytest tests/test_gpu_conv.py -m gpu -q -k "tcgen05 or tc_supported" > gpurun_out/t3.log 2>&1
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q > gpurun_out/t4.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
for f in t1 t2 t3 t4 smoke bench; do echo "=== $f"; tail -n 25 gpurun_out/$f.log; done
