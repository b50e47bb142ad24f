N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/check_strip_net.py > gpurun_out/r2z_strip_$N.log 2> gpurun_out/r2z_strip_$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 scripts/check_sharded_render.py > gpurun_out/r2z_shard_$N.log 2> gpurun_out/r2z_shard_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2z_bench_$N.log 2> gpurun_out/r2z_bench_$N.err
if [ "$N" = "2" ]; then timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2z_bench_1.log 2> gpurun_out/r2z_bench_1.err; fi
for f in r2z_strip_$N r2z_shard_$N r2z_bench_$N r2z_bench_1; do echo "=== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-2600; tail -n 5 gpurun_out/$f.err | grep -v "OMP_NUM\|\*\*\*\*" | cut -c1-300; done
