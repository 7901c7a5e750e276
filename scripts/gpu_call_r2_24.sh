mkdir -p gpurun_out
N=$1
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5) > gpurun_out/r2_24_bench_n$N.json 2> gpurun_out/r2_24_bench_n$N.err
tail -c 300 gpurun_out/r2_24_bench_n$N.err
