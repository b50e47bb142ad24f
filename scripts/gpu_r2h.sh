# usage: bash scripts/gpu_r2h.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/check_strip_net.py 256 256 300000 > gpurun_out/r2h_strip_small_$N.log 2> gpurun_out/r2h_strip_small_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 scripts/check_strip_net.py > gpurun_out/r2h_strip_$N.log 2> gpurun_out/r2h_strip_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2h_bench_$N.log 2> gpurun_out/r2h_bench_$N.err
timeout 300 python scripts/bench_raster_stream.py gpurun_out/r2h_raster.json > gpurun_out/r2h_raster.log 2>&1
for f in r2h_strip_small_$N r2h_strip_$N r2h_bench_$N; do echo "=== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-2500; tail -n 6 gpurun_out/$f.err | cut -c1-300; done
echo "=== raster"; tail -n 8 gpurun_out/r2h_raster.log
