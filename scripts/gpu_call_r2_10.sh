mkdir -p gpurun_out
free -g | head -2; nproc
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3) > gpurun_out/r2_10_bench_n8.json 2> gpurun_out/r2_10_bench_n8.err
tail -c 600 gpurun_out/r2_10_bench_n8.err
