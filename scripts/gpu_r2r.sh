mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1
echo "=== tests rc=$?"; tail -5 gpurun_out/r2r_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu --layer-times gpurun_out/r2r_layers.json > gpurun_out/r2r_bench.log 2>&1
echo "=== bench rc=$?"; tail -1 gpurun_out/r2r_bench.log | cut -c1-300
python scripts/show_layers.py gpurun_out/r2r_layers.json 2>&1 | head -90
export READ_B200_LIB=$PWD/read_b200/libread_b200_diag.so
timeout 300 python scripts/ab_pair_dbg.py "Encoder.1.layers.0.main.0,Encoder.1.layers.0.main.1" "tc_tma_store=0,tc_tma_store=1" > gpurun_out/r2r_ab_diag.log 2>&1
tail -3 gpurun_out/r2r_ab_diag.log
