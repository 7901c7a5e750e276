mkdir -p gpurun_out
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --big-scale 0.05) > gpurun_out/r2_03_bench_n2.json 2> gpurun_out/r2_03_bench_n2.err
tail -c 1500 gpurun_out/r2_03_bench_n2.err
python -m pytest tests -m gpu -x -q -k "multi_gpu" 2>&1 | tail -3
