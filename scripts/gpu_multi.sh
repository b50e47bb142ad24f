# usage: bash scripts/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m_bench_$N.log 2> gpurun_out/m_bench_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 scripts/check_sharded_render.py > gpurun_out/m_check_$N.log 2>&1
echo "=== bench $N"; tail -n 3 gpurun_out/m_bench_$N.log | cut -c1-2500; tail -n 5 gpurun_out/m_bench_$N.err
echo "=== check $N"; tail -n 5 gpurun_out/m_check_$N.log
