mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/r3l_tests.log 2>&1
echo "=== tests rc=$?"; tail -5 gpurun_out/r3l_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r3l_bench.log 2> gpurun_out/r3l_bench.err
echo "=== bench rc=$?"; tail -1 gpurun_out/r3l_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['reference_surface'])"
tail -2 gpurun_out/r3l_bench.err
