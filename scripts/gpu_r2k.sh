N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/check_dp_train.py > gpurun_out/r2k_dp_$N.log 2> gpurun_out/r2k_dp_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2k_bench_$N.log 2> gpurun_out/r2k_bench_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config c5 --steps 10 --warmup 3 > gpurun_out/r2k_bench_c5_$N.log 2> gpurun_out/r2k_bench_c5_$N.err
timeout 600 python scripts/ab_commit.py gpurun_out/r2k_ab.json > gpurun_out/r2k_ab.log 2>&1
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/r2k_t1.log 2>&1
for f in r2k_dp_$N r2k_bench_$N r2k_bench_c5_$N; do echo "=== $f"; tail -n 6 gpurun_out/$f.log | cut -c1-3000; tail -n 5 gpurun_out/$f.err | grep -v "OMP_NUM\|\*\*\*\*" | cut -c1-300; done
for f in r2k_ab r2k_t1; do echo "=== $f"; tail -n 16 gpurun_out/$f.log | cut -c1-300; done
