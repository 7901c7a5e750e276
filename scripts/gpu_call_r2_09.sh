mkdir -p gpurun_out
free -g | head -2; nproc
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/r2_09_bench_n2.json 2> gpurun_out/r2_09_bench_n2.err
tail -c 800 gpurun_out/r2_09_bench_n2.err
