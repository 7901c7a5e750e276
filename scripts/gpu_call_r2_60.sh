mkdir -p gpurun_out
(time python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2_60_bench_n1.json 2> gpurun_out/r2_60_bench_n1.err; tail -3 gpurun_out/r2_60_bench_n1.err
